#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k fused > gpurun_out/r2_n2_pytest_n$N.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_n2_pytest_n$N.log
run() { # name, extra env/args...
  name=$1; shift
  timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 20 $EXTRA > gpurun_out/r2_n2_${name}_n$N.log 2>&1
  echo "$name rc=$?"
  grep '^{' gpurun_out/r2_n2_${name}_n$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('push_update_phases_us'))"
}
EXTRA="" run default X=1
EXTRA="" run default_again X=1
