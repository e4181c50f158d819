#!/bin/bash
# Round-2 GPU session D: full parity suite + ncu evidence (launch lists, full-set captures of the four tensor-core kernels).
mkdir -p gpurun_out
# (parity suite: run by tools/gpu_final.sh)

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_render.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_under_ncu.txt 2>&1; echo "launch list (render) exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_train.csv \
    python tools/train_step.py 3 > gpurun_out/train_under_ncu.txt 2>&1; echo "launch list (train) exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_umma -s 2 -c 1 -o gpurun_out/r02_fused -f \
    python tools/prof_fused.py 307200 > gpurun_out/r02_fused.txt 2>&1; echo "full-set fused exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bwd_chain -s 3 -c 1 -o gpurun_out/r02_chain -f \
    python tools/train_step.py 3 > gpurun_out/r02_chain.txt 2>&1; echo "full-set chain exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_umma -s 3 -c 1 -o gpurun_out/r02_fwdtrain -f \
    python tools/train_step.py 3 > gpurun_out/r02_fwdtrain.txt 2>&1; echo "full-set training forward exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 40 -c 1 -o gpurun_out/r02_dw -f \
    python tools/train_step.py 3 > gpurun_out/r02_dw.txt 2>&1; echo "full-set dW exit $?"
ls -la gpurun_out/*.ncu-rep
