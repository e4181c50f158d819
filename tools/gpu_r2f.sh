#!/bin/bash
# round 2, late: batched dW launches + device-side Hungarian assignment -- training tests, per-kernel breakdown, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -15 gpurun_out/pytest_gpu.txt
HEADN=30 bash tools/gpu_variants.sh tools/prof_train.py dm-nerf_b200/lib/libdmnerf_b200.so
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_now.txt 2>&1
python -c "
import json
for l in open('gpurun_out/bench_now.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']); t=d['train_step']; print(t['ms_per_step'], t['roofline']['frac'], t['native_launches_per_step'], t['full_iteration'], t['penalizer'])
"
tail -3 gpurun_out/bench_now.txt | cut -c1-300
