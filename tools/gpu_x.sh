#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
for cfg in "xdeepfm 9" "dcn 64" "lr 9"; do set -- $cfg
timeout 300 python bench.py --model $1 --dim $2 --steps 200 --warmup 10 > gpurun_out/r2_x_$1.log 2>&1; echo "$1 rc=$?"; grep '^{' gpurun_out/r2_x_$1.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['final_loss'])" || tail -5 gpurun_out/r2_x_$1.log
done
