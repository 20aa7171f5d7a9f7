#!/bin/bash
# 1-GPU call: the whole GPU test suite, the model/dim sweep of BASELINE.md (ours + same-box NCCL/cuBLAS baseline),
# the host-tier cache-size sweep, and the default bench line. Results: gpurun_out/r2_sweep.json
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_full.log
one() { # tag impl model dim steps extra...
  tag=$1; impl=$2; model=$3; dim=$4; steps=$5; shift 5
  timeout 400 python bench.py --impl $impl --model $model --dim $dim --steps $steps --warmup 10 "$@" > gpurun_out/r2_sweep_${tag}.log 2>&1
  echo "$tag rc=$? $(grep '^{' gpurun_out/r2_sweep_${tag}.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3),'M/s', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e6,3))" 2>/dev/null)"
}
for cfg in "deepfm 64" "deepfm 9" "wdl 64" "wdl 9" "xdeepfm 9" "dcn 64" "lr 9"; do
  set -- $cfg
  one ours_$1_$2 ours $1 $2 200
  case $1 in deepfm|wdl) one base_$1_$2 baseline $1 $2 40;; esac   # the baseline arm covers DeepFM / WDL
done
one ours_deepfm_64_adam ours deepfm 64 200 --optimizer adam
one ours_deepfm_64_ftrl ours deepfm 64 200 --optimizer ftrl
for c in 262144 1048576 4194304; do
  timeout 400 python benchmarks/host_tier_bench.py --cache-rows $c --steps 200 --warmup 100 > gpurun_out/r2_tier_$c.log 2>&1
  echo "tier $c rc=$? $(grep '^{' gpurun_out/r2_tier_$c.log | tail -1 | cut -c1-400)"
done
timeout 600 python bench.py --impl reference > gpurun_out/r2_final_ref_n1.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_final_n1.log 2>&1; echo "bench rc=$?"; grep '^{' gpurun_out/r2_final_n1.log | tail -1 | cut -c1-300
python - <<'PY'
import glob, json, os
out = {}
for f in sorted(glob.glob("gpurun_out/r2_sweep_*.log")) + sorted(glob.glob("gpurun_out/r2_tier_*.log")):
    lines = [l for l in open(f) if l.startswith("{")]
    if lines:
        try:
            out[os.path.basename(f)[:-4]] = json.loads(lines[-1])
        except Exception as e:
            out[os.path.basename(f)[:-4]] = {"error": str(e)}
    else:
        out[os.path.basename(f)[:-4]] = {"error": open(f).read()[-600:]}
json.dump(out, open("gpurun_out/r2_sweep.json", "w"), indent=1)
PY
