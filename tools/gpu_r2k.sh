#!/bin/bash
# same-box A/B on the training step: store cache operators of the plane stores, 256-bit loads in the dW loader
HEADN=14 bash tools/gpu_variants.sh tools/prof_train.py dm-nerf_b200/lib/libdmnerf_b200.so tools/bin/v_stcs.so tools/bin/v_stcg.so tools/bin/v_ld256.so dm-nerf_b200/lib/libdmnerf_b200.so 2>&1 | grep -v "autograd::engine\|Optimizer"
