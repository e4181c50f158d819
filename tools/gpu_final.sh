#!/bin/bash
# Round-2 final GPU session: full parity suite, smoke, stress, bench lines (three workloads), refreshed ncu evidence.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -2 gpurun_out/smoke.txt
timeout 300 python tools/stress_train.py 300 2>&1 | grep -E "STATUS|slow steps"
timeout 200 python tools/stress_render.py 150 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.txt 2>&1; echo "bench exit $?"
timeout 600 python bench.py --steps 5 --warmup 3 --workload replica_room0 --no-train --no-cpu-baseline > gpurun_out/bench_room0.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload replica_room0_93 --no-train --no-cpu-baseline > gpurun_out/bench_room0_93.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload replica_office2 --no-train --no-cpu-baseline > gpurun_out/bench_office2.txt 2>&1
bash tools/gpu_r2d.sh > gpurun_out/r2d.log 2>&1; tail -8 gpurun_out/r2d.log
python - <<'PY'
import json
for f in ("bench_default", "bench_room0", "bench_room0_93", "bench_office2"):
    for l in open("gpurun_out/%s.txt" % f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.0f rays/s" % d["value"], "%.2f ms" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], d["clocks"]["sm_mhz"])
            if "train_step" in d:
                t = d["train_step"]; print("  train %.3f ms frac %.4f full %.3f ms pen %.1f us" % (t["ms_per_step"], t["roofline"]["frac"], t["full_iteration"]["ms_per_step"], t["penalizer"]["us_per_fwd_bwd"]))
PY
