#!/bin/bash
# Round 2, GPU call 4 (1 x B200): verification after the code-size fix + endomorphism subgroup check, then the
# round-2 ncu --set full capture of k_msm_affine (the source of profiles/roofline_inputs.json).
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_verify.py tests/test_gpu_callers.py -m gpu -x -q ) > gpurun_out/r02_c4_pytest_verify.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c4_pytest_verify.txt
tail -4 gpurun_out/r02_c4_pytest_verify.txt
RAIKO_KZG_VERIFY_TRACE=1 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c4_verify_lanes.txt 2>&1
tail -8 gpurun_out/r02_c4_verify_lanes.txt
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
$CMD > gpurun_out/r02_c4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_msm_affine -s 2 -c 1 -o gpurun_out/prof_msm_affine_r02 $CMD > gpurun_out/r02_c4_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -2
