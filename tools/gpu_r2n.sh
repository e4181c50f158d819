#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_fuzz.py tests/test_gpu_edge.py -m gpu -q -x 2>&1 | tail -3
HEADN=16 bash tools/gpu_variants.sh tools/prof_train.py dm-nerf_b200/lib/libdmnerf_b200.so 2>&1 | grep "composite\|device time\|Backward"
