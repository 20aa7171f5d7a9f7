#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 $EXTRA > gpurun_out/r2_bench_${name}_n$N.log 2>&1
  echo "$name rc=$?"
  grep '^{' gpurun_out/r2_bench_${name}_n$N.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,2),'M/s', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d.get('push_update_phases_us'), d.get('sparse_counters'))"
}
EXTRA="--prefetch" run v2pf EXB_SPARSE_V2=1
EXTRA="" run v1 EXB_SPARSE_V2=0
EXTRA="" run v2 EXB_SPARSE_V2=1
