#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_umma -s 2 -c 1 -o gpurun_out/prof_fused -f \
    python tools/prof_fused.py > gpurun_out/prof_fused.txt 2>&1
echo "full-set exit $?"; tail -3 gpurun_out/prof_fused.txt
