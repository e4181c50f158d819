#!/bin/bash
# Exactly what the driver runs at round end: GPU tests, smoke, default bench, reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 900 python bench.py > gpurun_out/bench_default.txt 2>&1; echo "bench exit $?" >> gpurun_out/bench_default.txt
tail -n 6 gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt; tail -n 2 gpurun_out/bench_default.txt | cut -c1-2500
