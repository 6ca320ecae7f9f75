#!/bin/bash
# Round 2, GPU call 18 (1 x B200): lockstep groups x hot-split form around the Karatsuba kernel.
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base f4s4 f4s8 f3s8 f4s0 2>&1 | tee gpurun_out/r02_c18_ab.txt
