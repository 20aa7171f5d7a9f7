#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -k "cin or tc_linear" 2>&1 | tail -15
timeout 300 python bench.py --model xdeepfm --dim 9 --steps 100 --warmup 10 > gpurun_out/r2_x_xdeepfm.log 2>&1; echo "xdeepfm rc=$?"; grep '^{' gpurun_out/r2_x_xdeepfm.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['final_loss'])" || tail -5 gpurun_out/r2_x_xdeepfm.log
