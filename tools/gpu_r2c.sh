#!/bin/bash
# Round-2 GPU session C: split-chunk epilogue -- parity, A/B against the contiguous-column build, in-kernel cycle profile.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
bash tools/gpu_ab.sh tools/bin/v_nosplit.so dm-nerf_b200/lib/libdmnerf_b200.so
timeout 300 python tools/kprof.py --fused > gpurun_out/kprof_fused_split.txt 2>&1
KPROF_LIB=tools/bin/libdmnerf_kprof_dmn_epi_split=0.so timeout 300 python tools/kprof.py --fused > gpurun_out/kprof_fused_nosplit.txt 2>&1
head -18 gpurun_out/kprof_fused_split.txt; head -18 gpurun_out/kprof_fused_nosplit.txt
