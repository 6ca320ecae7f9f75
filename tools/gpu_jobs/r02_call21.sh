#!/bin/bash
# Round 2, GPU call 21 (1 x B200): cp.async staging of the first-needed operands of k_msm_affine's forward /
# backward iterations (RK_AFF_STAGE=1) against the shipped kernel.
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base stfwd stbwd dummy 2>&1 | tee gpurun_out/r02_c21_ab.txt
