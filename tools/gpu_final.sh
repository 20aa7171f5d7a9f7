#!/bin/bash
# final 1-GPU check of the tree: whole GPU suite, smoke(), default bench + reference arm, the eager zoo numbers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$? $(grep -c SMOKE_OK gpurun_out/r2_final_smoke.log)"
timeout 300 python bench.py --impl reference > gpurun_out/r2_final_ref_n1.log 2>&1; tail -1 gpurun_out/r2_final_ref_n1.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2_final_n1.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2_final_n1.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
for cfg in "xdeepfm 9" "dcn 64"; do set -- $cfg
timeout 300 python bench.py --model $1 --dim $2 --steps 200 --warmup 10 > gpurun_out/r2_final_$1.log 2>&1; echo "$1 rc=$?"; grep '^{' gpurun_out/r2_final_$1.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"
done
