#!/bin/bash
# Round 2, GPU call 6 (1 x B200): A/B of k_msm_affine builds -- Montgomery-product formulations (RK_MUL_FORM 0/1/2),
# fused limb passes (RK_AFF_FUSE), 20 / 24 warps per SM at 96 / 80 registers -- and the L2 fetch-granularity knob.
set -u
mkdir -p gpurun_out
python tests/tools/gpu_lib_ab.py base f1 f2 f0u f1u f2w20:20 f2w20u:20 f1w20:20 f1w20u:20 f2w24:24 2>&1 | tee gpurun_out/r02_c6_ab.txt
RAIKO_KZG_L2_FETCH=32 python tests/tools/gpu_lib_ab.py base f1u 2>&1 | sed 's/^/L2_FETCH=32 /' | tee -a gpurun_out/r02_c6_ab.txt
