#!/bin/bash
# GPU call 1 (round 2): existing suite on this round's box, ours vs baseline at N=1, smoke, sanitizers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke1.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke1.log
timeout 400 python bench.py --steps 300 --warmup 20 > gpurun_out/r2_bench_ours1.log 2>&1; echo "ours rc=$?"; tail -1 gpurun_out/r2_bench_ours1.log | cut -c1-400
timeout 400 python bench.py --impl baseline --steps 100 --warmup 10 > gpurun_out/r2_bench_base_bf16.log 2>&1; echo "base rc=$?"; tail -1 gpurun_out/r2_bench_base_bf16.log | cut -c1-400
timeout 400 python bench.py --impl baseline --baseline-dtype fp32 --steps 100 --warmup 10 > gpurun_out/r2_bench_base_fp32.log 2>&1; echo "base32 rc=$?"; tail -1 gpurun_out/r2_bench_base_fp32.log | cut -c1-300
SAN_TIMEOUT=240 bash tools/sanitize_gpu.sh
