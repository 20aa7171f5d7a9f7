#!/bin/bash
# N-GPU call (N = $1): real multi-GPU checks + the default bench line + the same-box baseline at this N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/mp_gpu_check.py > gpurun_out/r2_final_mpcheck_n$N.log 2>&1; echo "mp_check rc=$? $(grep -c MP_GPU_CHECK_PASSED gpurun_out/r2_final_mpcheck_n$N.log)"
EXB_TEST_PREFETCH=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tests/mp_gpu_fused_check.py > gpurun_out/r2_final_mpfused_n$N.log 2>&1; echo "mp_fused rc=$? $(grep MP_GPU_FUSED_PASSED gpurun_out/r2_final_mpfused_n$N.log)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 20 > gpurun_out/r2_final_ours_n$N.log 2>&1; echo "ours rc=$?"
grep '^{' gpurun_out/r2_final_ours_n$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('push_update_phases_us'))"
if [ "$2" = "base" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl baseline --steps 100 --warmup 10 > gpurun_out/r2_final_base_n$N.log 2>&1; echo "base rc=$?"
grep '^{' gpurun_out/r2_final_base_n$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
fi
