#!/bin/bash
# Round 2, GPU call 14 (1 x B200): one-level subtractive Karatsuba on the product rows (RK_MUL_FORM=3): register-resident
# chain rate and k_msm_affine A/B at 16 x 128 and 12 x 168 registers.
set -u
mkdir -p gpurun_out
for f in 0 3; do echo "== RK_MUL_FORM=$f"; tools/ubench/fpmul_ubench_form$f | grep -E "mul chain|2 mul chains"; done 2>&1 | tee gpurun_out/r02_c14_fpmul_forms.txt
python tests/tools/gpu_lib_ab.py base f3 f3w12:12 f0w12:12 2>&1 | tee gpurun_out/r02_c14_ab.txt
