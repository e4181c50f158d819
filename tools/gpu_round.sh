#!/bin/bash
# One GPU session: hardware probe, GPU parity tests, smoke, short bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc >> gpurun_out/nvsmi.txt
timeout 300 python tools/umma_check.py > gpurun_out/umma_check.txt 2>&1; echo "umma_check exit $?" >> gpurun_out/umma_check.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 1200 python bench.py --steps 2 --warmup 3 ${BENCH_ARGS:---mlp auto} > gpurun_out/bench.txt 2>&1; echo "bench exit $?" >> gpurun_out/bench.txt
tail -n 30 gpurun_out/umma_check.txt gpurun_out/pytest_gpu.txt gpurun_out/smoke.txt gpurun_out/bench.txt
