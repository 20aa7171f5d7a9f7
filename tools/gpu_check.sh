#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/pytest_gpu14.log 2>&1; tail -3 gpurun_out/pytest_gpu14.log
EXB_GEMM_BN=64 timeout 300 python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/bench_v14.log
python -c "import json; d=json.loads(open('gpurun_out/bench_v14.log').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'])"
EXB_GEMM_BN=64 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches14.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench14.log 2>&1; python tools/step_timeline.py gpurun_out/launches14.csv | grep -v gemm | tail -8
