#!/bin/bash
# Round 2, GPU call 25 (1 x B200): 14 warps x 144 registers and 18 warps x 112 registers around the shipped 16 x 128.
set -u
mkdir -p gpurun_out
AB_REPS=3 python tests/tools/gpu_lib_ab.py base w14s2:14 w14s7:14 w18s3:18 w18s6:18 2>&1 | tee gpurun_out/r02_c25_ab.txt
