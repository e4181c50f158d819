#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3
HEADN=16 bash tools/gpu_variants.sh tools/prof_train.py dm-nerf_b200/lib/libdmnerf_b200.so 2>&1 | grep -v "autograd::engine\|Optimizer"
DMNERF_SELECT=device PARTS=select,ins,pen TOPN=0 timeout 200 python tools/prof_full_iter.py 2>/dev/null
ADAM=fused DMNERF_SELECT=device PARTS=select,ins,pen TOPN=0 timeout 200 python tools/prof_full_iter.py 2>/dev/null
