#!/bin/bash
# Round 2, GPU call 23 (1 x B200): 16-bit digit codes (RK_AFF_CODE16) with 64 / 96 / 128 chains per lane against the
# shipped kernel (32-bit codes, 64 chains).
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base 2>&1 | tee gpurun_out/r02_c23_ab.txt
for k in 64 96 128; do RAIKO_KZG_AFFINE_CHAINS=$k AB_REPS=3 python tests/tools/gpu_lib_ab.py c16 2>&1 | sed "s/^/K=$k /" | tee -a gpurun_out/r02_c23_ab.txt; done
