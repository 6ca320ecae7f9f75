#!/bin/bash
# Round 2, GPU call 11 (1 x B200): ncu --set full of the CTA-wide pairing check (where do its 9 ms go?)
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_pairing_check_cta -s 1 -c 1 -o gpurun_out/prof_pairing_cta_r02 python tests/tools/verify_trace.py 256 > gpurun_out/r02_c11_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_pairing_cta_r02.ncu-rep
