#!/bin/bash
# Round 2, GPU call 17 (1 x B200): A/B around the Karatsuba kernel (base = RK_MUL_FORM 3): hot split at bit 32 (f4),
# Karatsuba square (f5), 20 warps x 96 registers, 4 / 1 lockstep groups, unfused limb passes.
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base f4 f5 f3w20:20 f3s4 f3s1 f3nf base 2>&1 | tee gpurun_out/r02_c17_ab.txt
