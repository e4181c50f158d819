#!/bin/bash
# Build the working tree's library with extra nvcc flags into tools/bin/v_<name>.so (same-box A/B runs:
# DMNERF_LIB_PATH=tools/bin/v_<name>.so python ...).     tools/variant_build.sh <name> "<extra flags>"
set -e
name=$1; shift
mkdir -p tools/bin /tmp/v_$name
objs=""
for f in dm-nerf_b200/csrc/*.cu; do
  o=/tmp/v_$name/$(basename ${f%.cu}).o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $@ -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/bin/v_$name.so $objs -lcudart
echo tools/bin/v_$name.so
