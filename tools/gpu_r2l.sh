#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -2 gpurun_out/smoke.txt
