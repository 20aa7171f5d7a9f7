#!/bin/bash
timeout 600 python -m pytest tests/ -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; tail -3 gpurun_out/pytest_gpu17.log
timeout 200 python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/bench_v17.log
python -c "import json; d=json.loads(open('gpurun_out/bench_v17.log').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'])"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -s 30 -c 9 -f -o gpurun_out/prof_gemm python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/gemm_raw.csv 2>/dev/null; ls -la gpurun_out/prof_gemm.ncu-rep | cut -c1-80
