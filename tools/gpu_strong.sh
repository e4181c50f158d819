#!/bin/bash
# BASELINE configs[4]: strong scaling of ONE office_2 frame over N GPUs (pixel ranges, device-side ray generation, one all-gather).
mkdir -p gpurun_out
: > gpurun_out/strong_scaling.txt
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --workload replica_office2 --scaling strong --no-train --no-cpu-baseline >> gpurun_out/strong_scaling.txt 2>gpurun_out/strong_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 8 --warmup 3 --workload replica_office2 --scaling strong --no-train --no-cpu-baseline >> gpurun_out/strong_scaling.txt 2>gpurun_out/strong_$n.err
  fi
  echo "rc $?" >> gpurun_out/strong_scaling.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/weak8.txt 2>gpurun_out/weak8.err; echo "rc $?" >> gpurun_out/weak8.txt
python - <<'PY'
import json
for f in ("gpurun_out/strong_scaling.txt", "gpurun_out/weak8.txt"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["n_gpus"], d["scaling"], "ms/step %.2f" % d["ms_per_step"], "rays/s %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], d.get("allgather"), d["clocks"]["sm_mhz"])
        elif l.startswith("rc"):
            print(l.strip())
PY
