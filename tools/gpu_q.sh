#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sparse_engine.py tests/test_gpu_fused.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r2_q.log 2>&1; echo rc=$?
grep '^{' gpurun_out/r2_q.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['push_update_phases_us'], d.get('pull_probe_us'))"
EXB_SPARSE_V2=1 timeout 300 python tools/mp_timeline.py --steps 40 2>&1 | grep -E "^rank 0|phases"
