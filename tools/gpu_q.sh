#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_fused.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r2_q.log 2>&1; echo rc=$?
grep '^{' gpurun_out/r2_q.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"
timeout 600 python benchmarks/checkpoint_bench.py --rows 8000000 > gpurun_out/r2_ckpt_bench.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_ckpt_bench.log
