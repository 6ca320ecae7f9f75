#!/bin/bash
# Round 2, GPU call 29 (1 x B200): k_fr_eval_quot at 3 and 4 resident CTAs per SM (80 / 64 registers) against 2 (128).
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base fr3 fr4 2>&1 | tee gpurun_out/r02_c29_ab.txt
