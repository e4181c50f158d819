#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -14 gpurun_out/smoke.txt
HEADN=30 bash tools/gpu_variants.sh tools/prof_train.py dm-nerf_b200/lib/libdmnerf_b200.so
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_now.txt 2>&1
python -c "
import json
for l in open('gpurun_out/bench_now.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']); t=d['train_step']; print(t['ms_per_step'], t['roofline']['frac'], t['full_iteration'], t['penalizer'], t.get('cpu_reference')); print(d['cpu_baseline'])
"
