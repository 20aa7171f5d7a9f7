#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -12
for pf in "" "--no-prefetch"; do
timeout 400 python bench.py --steps 300 --warmup 20 $pf > gpurun_out/r2_q_pf.log 2>&1; echo "pf='$pf' rc=$?"
grep '^{' gpurun_out/r2_q_pf.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d['e2e_final_loss'])"
done
