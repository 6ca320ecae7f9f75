#!/bin/bash
# Round 2, GPU call 22 (1 x B200): cache hints for the chain scratch (streaming prefix products / all stores) and a
# chains-per-lane sweep with four lockstep groups.
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base ch1 ch2 2>&1 | tee gpurun_out/r02_c22_ab.txt
for k in 32 48 56; do RAIKO_KZG_AFFINE_CHAINS=$k AB_REPS=2 python tests/tools/gpu_lib_ab.py base 2>&1 | sed "s/^/K=$k /" | tee -a gpurun_out/r02_c22_ab.txt; done
