#!/bin/bash
# Dev iteration: UMMA check + GPU tests + short bench without the CPU arm.
mkdir -p gpurun_out
timeout 300 python tools/umma_check.py > gpurun_out/umma_check.txt 2>&1; echo "umma_check exit $?" >> gpurun_out/umma_check.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.txt 2>&1; echo "bench exit $?" >> gpurun_out/bench.txt
tail -n 12 gpurun_out/umma_check.txt gpurun_out/pytest_gpu.txt gpurun_out/bench.txt
