#!/bin/bash
# Round-2 GPU session A: parity tests, parity-at-scale table, training-step kernel breakdown, bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
nproc >> gpurun_out/nvsmi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 600 python tools/parity_at_scale.py 2048 > gpurun_out/parity_at_scale.txt 2>&1; echo "parity exit $?" >> gpurun_out/parity_at_scale.txt
timeout 300 python tools/prof_train.py > gpurun_out/prof_train.txt 2>&1; echo "prof_train exit $?" >> gpurun_out/prof_train.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.txt 2>&1; echo "bench exit $?" >> gpurun_out/bench_default.txt
timeout 600 python bench.py --steps 3 --warmup 3 --workload replica_room0 --no-train --no-cpu-baseline > gpurun_out/bench_room0.txt 2>&1; echo "bench exit $?" >> gpurun_out/bench_room0.txt
timeout 600 python bench.py --steps 3 --warmup 3 --workload replica_office2 --scaling strong --no-train --no-cpu-baseline > gpurun_out/bench_office2_strong1.txt 2>&1; echo "bench exit $?" >> gpurun_out/bench_office2_strong1.txt
tail -n 25 gpurun_out/pytest_gpu.txt gpurun_out/parity_at_scale.txt gpurun_out/prof_train.txt
tail -c 1500 gpurun_out/bench_default.txt; tail -c 600 gpurun_out/bench_room0.txt gpurun_out/bench_office2_strong1.txt
