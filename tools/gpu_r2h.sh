#!/bin/bash
# full validation of the late round-2 tree: every GPU test, smoke, training stress, per-kernel breakdown, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt; tail -4 gpurun_out/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt; tail -3 gpurun_out/smoke.txt
timeout 300 python tools/stress_train.py 300 > gpurun_out/stress.txt 2>&1; tail -2 gpurun_out/stress.txt
HEADN=40 bash tools/gpu_variants.sh tools/prof_train.py dm-nerf_b200/lib/libdmnerf_b200.so > /dev/null; cp gpurun_out/variants.txt gpurun_out/prof_train.txt; head -24 gpurun_out/prof_train.txt
PARTS=select,ins,pen TOPN=0 timeout 200 python tools/prof_full_iter.py 2>/dev/null | tee gpurun_out/full_iter_prof.txt
DMNERF_SELECT=device PARTS=select,ins,pen TOPN=0 timeout 200 python tools/prof_full_iter.py 2>/dev/null | tee -a gpurun_out/full_iter_prof.txt
timeout 600 python bench.py > gpurun_out/bench_now.txt 2>&1
python -c "
import json
for l in open('gpurun_out/bench_now.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']); t=d['train_step']; print(t['ms_per_step'], t['roofline']['frac'], t['native_launches_per_step'], t['full_iteration']['ms_per_step'], t.get('full_iteration_device_select'), t['penalizer']['us_per_fwd_bwd'])
"
