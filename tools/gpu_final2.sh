#!/bin/bash
# last validation of the round-2 tree: every GPU test, smoke, the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -2 gpurun_out/smoke.txt
timeout 600 python bench.py > gpurun_out/bench_now.txt 2>&1
python -c "
import json
for l in open('gpurun_out/bench_now.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); t=d['train_step']; print(t['ms_per_step'], t['roofline']['frac'], t['native_launches_per_step'], t['full_iteration']['ms_per_step'], t['full_iteration_device_select']['ms_per_step'], t['penalizer']['us_per_fwd_bwd'])
"
