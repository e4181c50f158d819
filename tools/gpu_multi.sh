#!/bin/bash
# Multi-GPU bench as the driver launches it (torchrun, one rank per GPU, NCCL) + the reference (CPU) arm.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_gpus$N.txt 2>&1; echo "bench x$N exit $?" >> gpurun_out/bench_gpus$N.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_reference.txt 2>&1; echo "ref exit $?" >> gpurun_out/bench_reference.txt
tail -n 5 gpurun_out/bench_gpus$N.txt gpurun_out/bench_reference.txt | cut -c1-1500
