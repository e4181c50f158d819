#!/bin/bash
# Tensor-core backward: parity tests + A/B timing of the training step against the fp32 CUDA-core GEMMs.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/bwd_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/bwd_tests.txt
tail -n 15 gpurun_out/bwd_tests.txt
for impl in tc simt; do
  DMNERF_BWD_IMPL=$impl timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$impl', 'train step ms', d['train_step']['ms_per_step'], 'rays/s', d['train_step']['rays_per_s'])
"
done
