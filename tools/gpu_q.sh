#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/loss_parity.py --steps 2000 --dim 16 > gpurun_out/r2_loss_parity.log 2>&1; echo rc=$?
tail -1 gpurun_out/r2_loss_parity.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: d[k] for k in ('first_window','last_window','fused_vs_fp32','bf16_baseline_vs_fp32')})" || tail -5 gpurun_out/r2_loss_parity.log
