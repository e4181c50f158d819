#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge.py tests/test_gpu_stages.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-train > gpurun_out/bench_now.txt 2>&1
python -c "
import json
for l in open('gpurun_out/bench_now.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'], d['clocks'])
"
tail -2 gpurun_out/bench_now.txt | cut -c1-200
