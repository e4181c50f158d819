#!/bin/bash
# Round-2 GPU session B: training path with the fused gradient chain.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_scale.py -m gpu -q -x --durations=5 > gpurun_out/pytest_train.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.txt
timeout 300 python tools/prof_train.py > gpurun_out/prof_train_chain.txt 2>&1; echo "exit $?" >> gpurun_out/prof_train_chain.txt
DMNERF_BWD_IMPL=gemm timeout 300 python tools/prof_train.py > gpurun_out/prof_train_gemm.txt 2>&1; echo "exit $?" >> gpurun_out/prof_train_gemm.txt
tail -n 30 gpurun_out/pytest_train.txt; head -34 gpurun_out/prof_train_chain.txt; head -8 gpurun_out/prof_train_gemm.txt
