#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_gpus$N.txt 2>&1; echo "bench x$N exit $?" >> gpurun_out/bench_gpus$N.txt
tail -n 3 gpurun_out/bench_gpus$N.txt | cut -c1-700
