#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sparse_engine.py tests/test_gpu_fused.py tests/test_gpu_model.py -x -q > gpurun_out/pytest_gpu15.log 2>&1; tail -3 gpurun_out/pytest_gpu15.log
python tools/sparse_probe.py --vocab 1tb --iters 4 2>&1 | tail -9
timeout 300 python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/bench_v15.log
python -c "import json; d=json.loads(open('gpurun_out/bench_v15.log').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d.get('push_update_phases_us'))"
