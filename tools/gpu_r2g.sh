#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge.py -m gpu -q -x -k round2 2>&1 | tail -3
for p in select,ins,pen ins,pen select,pen select,ins none; do PARTS=$p TOPN=$([ $p = select,ins,pen ] && echo 40 || echo 0) timeout 200 python tools/prof_full_iter.py 2>/dev/null; done | tee gpurun_out/full_iter_prof.txt
