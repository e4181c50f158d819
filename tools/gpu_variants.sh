#!/bin/bash
# Same-box timing of library variants on the training step: tools/gpu_variants.sh <script.py> lib1.so lib2.so ...
mkdir -p gpurun_out
script=$1; shift
: > gpurun_out/variants.txt
for lib in "$@"; do
  echo "== $lib" >> gpurun_out/variants.txt
  DMNERF_LIB_PATH=$lib timeout 300 python $script 2>/dev/null | head -${HEADN:-12} >> gpurun_out/variants.txt
done
cat gpurun_out/variants.txt
