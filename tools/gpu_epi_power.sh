#!/bin/bash
# Sustained rate of the bare network kernel with the hidden-step epilogue ablated (timing only; results are garbage):
# where does the ~1 kW go?  base / no bias-ReLU-split arithmetic / no TMEM traffic at all.
mkdir -p gpurun_out
python -c "import torch" 2>/dev/null
for v in notmem base notmem; do
  echo "== $v" >> gpurun_out/epi_power.log
  ( while true; do nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader; sleep 0.5; done ) > gpurun_out/epi_power_$v.smi &
  smi=$!
  DMNERF_LIB_PATH=tools/bin/v_$v.so timeout 120 python tools/raw_mlp_rate.py >> gpurun_out/epi_power.log 2>&1
  kill $smi
  sort -t, -k2 -n -r gpurun_out/epi_power_$v.smi | head -3 >> gpurun_out/epi_power.log
done
cat gpurun_out/epi_power.log
