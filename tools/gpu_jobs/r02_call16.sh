#!/bin/bash
# Round 2, GPU call 16 (1 x B200): final captures with the Karatsuba product rows: ncu launch list of a bench step,
# ncu --set full of k_msm_affine, k_fr_eval_quot (affine path) and of k_msm (XYZZ path forced) -- the sources of
# profiles/roofline_inputs.json.
set -u
mkdir -p gpurun_out
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
$CMD > gpurun_out/r02_c16_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c16_launches.csv $CMD > gpurun_out/r02_c16_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_msm_affine -s 2 -c 1 -o gpurun_out/prof_msm_affine_r02_final $CMD > gpurun_out/r02_c16_ncu_aff.log 2>&1; echo "ncu affine rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fr_eval_quot -s 1 -c 1 -o gpurun_out/prof_fr_r02_final $CMD > gpurun_out/r02_c16_ncu_fr.log 2>&1; echo "ncu fr rc=$?"
CMDX="python bench.py --batch 2368 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
RAIKO_KZG_MSM_AFFINE=0 ncu --set full --clock-control none --import-source on -k regex:k_msm$ -s 2 -c 1 -o gpurun_out/prof_msm_xyzz_r02_final $CMDX > gpurun_out/r02_c16_ncu_xyzz.log 2>&1; echo "ncu xyzz rc=$?"
ls -la gpurun_out/*_final.ncu-rep
