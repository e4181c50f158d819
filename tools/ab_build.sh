#!/bin/bash
# Build the product library of given commits into tools/bin/ab_<commit>.so for same-box A/B timing
# (run:  DMNERF_LIB_PATH=tools/bin/ab_<commit>.so python bench.py ...).
set -e
mkdir -p tools/bin
for c in "$@"; do
  d=/tmp/ab_$c; rm -rf $d; mkdir -p $d
  git archive $c dm-nerf_b200/csrc include | tar -x -C $d
  objs=""
  for f in $d/dm-nerf_b200/csrc/*.cu; do
    o=${f%.cu}.o
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -I$d/include -c $f -o $o &
    objs="$objs $o"
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/ab_$c.so $objs -lcudart
  echo tools/bin/ab_$c.so
done
