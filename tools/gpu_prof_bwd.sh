#!/bin/bash
# ncu --set full captures of the tensor-core backward GEMM kernels inside a training step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc_kernel -s 20 -c 1 -o gpurun_out/prof_bwd_tn -f \
    python tools/prof_train.py > gpurun_out/prof_bwd_tn.txt 2>&1
echo "tn exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nn_tc_kernel -s 20 -c 1 -o gpurun_out/prof_bwd_nn -f \
    python tools/prof_train.py > gpurun_out/prof_bwd_nn.txt 2>&1
echo "nn exit $?"
ls -la gpurun_out/*.ncu-rep
