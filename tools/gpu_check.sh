#!/bin/bash
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/pytest_gpu16.log 2>&1; tail -3 gpurun_out/pytest_gpu16.log
timeout 300 python bench.py --steps 200 --warmup 20 2>&1 | tail -1 > gpurun_out/bench_v16.log
python -c "import json; d=json.loads(open('gpurun_out/bench_v16.log').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['final_loss'], d.get('push_update_phases_us'), d['gpu_launches'], d['clocks'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
