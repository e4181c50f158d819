#!/bin/bash
# Same-box A/B of product-library builds: tools/gpu_ab.sh tools/bin/ab_x.so tools/bin/ab_y.so ...
mkdir -p gpurun_out
: > gpurun_out/${AB_OUT:-ab}.txt
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib (pass $rep)" >> gpurun_out/${AB_OUT:-ab}.txt
    DMNERF_LIB_PATH=$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-train ${AB_ARGS} 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('fused ms/frame %.2f  rays/s %.0f  frac %.4f  sm_mhz %s  power %s' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks'].get('power_w_max')))
" >> gpurun_out/${AB_OUT:-ab}.txt
  done
done
cat gpurun_out/${AB_OUT:-ab}.txt
