#!/bin/bash
# ncu evidence for profiles/: (1) launch list of bench steps, (2) full-set capture of one fused-render launch of the bench
# (one 640x480 frame), (3) full-set capture of the unfused fine-network kernel for comparison.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
echo "launch-list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mlp_umma -s 3 -c 1 -o gpurun_out/prof_fused -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_fused.txt 2>&1
echo "full-set (fused, one frame) exit $?"
timeout 900 ncu --set full --clock-control none -k regex:mlp_umma -s 2 -c 1 -o gpurun_out/prof_umma -f \
    python tools/prof_one.py > gpurun_out/prof_one.txt 2>&1
echo "full-set (unfused fine network) exit $?"
ls -la gpurun_out/ | head -30
