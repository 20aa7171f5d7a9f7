#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_engine.py tests/test_gpu_fused.py tests/test_gpu_model.py -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest2.log
EXB_PULL2=1 timeout 900 python -m pytest tests/test_gpu_sparse_engine.py -x -q -k "planned or virtual" > gpurun_out/r2_pytest2b.log 2>&1; echo "pytest pull2 rc=$?"
tail -3 gpurun_out/r2_pytest2b.log
b() { name=$1; shift
 timeout 400 env "$@" python bench.py --steps 300 --warmup 20 $EXTRA > gpurun_out/r2_bench_$name.log 2>&1; echo "$name rc=$?"
 grep '^{' gpurun_out/r2_bench_$name.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['push_update_phases_us'])"
}
EXTRA="" b ours_v2 EXB_SPARSE_V2=1
EXTRA="" b ours_v2_nopack EXB_SPARSE_V2=1 EXB_PACK_LINEAR=0
EXTRA="" b ours_v1 EXB_SPARSE_V2=0
EXB_SPARSE_V2=1 timeout 300 python tools/mp_timeline.py --steps 40 2>&1 | grep -E "^rank 0|phases"
