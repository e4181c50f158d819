#!/bin/bash
# ncu evidence: (1) launch list of a bench step, (2) full-set capture of the fine-network tcgen05 kernel.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
echo "launch-list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_umma -s 2 -c 1 -o gpurun_out/prof_umma -f \
    python tools/prof_one.py > gpurun_out/prof_one.txt 2>&1
echo "full-set exit $?"
ls -la gpurun_out/
