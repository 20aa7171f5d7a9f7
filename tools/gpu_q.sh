#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py -x -q 2>&1 | tail -8
for opt in adagrad adam ftrl; do
timeout 400 python bench.py --steps 200 --warmup 20 --optimizer $opt > gpurun_out/r2_q_opt_$opt.log 2>&1; echo "opt=$opt rc=$?"
grep '^{' gpurun_out/r2_q_opt_$opt.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d['config']['model'])"
done
