#!/bin/bash
# late round 2: refresh the ncu evidence of the training step (batched dW launches) and the sanitizer logs
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_train.csv \
    python tools/train_step.py 3 > gpurun_out/train_under_ncu.txt 2>&1; echo "launch list (train) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 6 -c 1 -o gpurun_out/r02_dw -f \
    python tools/train_step.py 3 > gpurun_out/r02_dw.txt 2>&1; echo "full-set dW (first wide batch of step 2) exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 9 -c 1 -o gpurun_out/r02_dw_b -f \
    python tools/train_step.py 3 > gpurun_out/r02_dw_b.txt 2>&1; echo "full-set dW (second wide batch of step 2) exit $?"
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.txt
SANITIZE_PARTS=0 timeout 500 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/sanitizer_synccheck.txt 2>&1; echo "synccheck exit $?"; tail -4 gpurun_out/sanitizer_synccheck.txt
ls -la gpurun_out/*.ncu-rep
