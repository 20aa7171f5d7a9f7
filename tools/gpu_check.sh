#!/bin/bash
# one-GPU regression + A/B bench (run through gpurun): GEMM tests, full GPU suite, bench configs
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm2.log 2>&1; tail -5 gpurun_out/pytest_gemm2.log
EXB_MN_MAJOR=0 timeout 600 python -m pytest tests/ -m gpu -x -q --deselect tests/test_gpu_gemm.py::test_dw_mn_major > gpurun_out/pytest_gpu9.log 2>&1; tail -3 gpurun_out/pytest_gpu9.log
for cfg in "0 0" "1 0" "1 1"; do set -- $cfg
  EXB_PDL=$1 EXB_MN_MAJOR=$2 timeout 300 python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/bench_cfg$1$2.log
  python -c "import json; d=json.loads(open('gpurun_out/bench_cfg$1$2.log').read()); print('pdl=$1 mn=$2', d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d.get('push_update_phases_us'))"
done
