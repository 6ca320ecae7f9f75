#!/bin/bash
# Round 2, GPU call 8 (1 x B200): lone-warp issue-rate microbenchmark; the default bench line (65,536 blobs) with the
# fused limb passes; ncu launch list and ncu --set full capture of the final k_msm_affine (source of
# profiles/roofline_inputs.json).
set -u
mkdir -p gpurun_out
tools/ubench/lone_warp_issue > gpurun_out/r02_c8_lone_warp_issue.txt 2>&1; cat gpurun_out/r02_c8_lone_warp_issue.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_c8_bench.json 2> gpurun_out/r02_c8_bench.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_c8_bench.json
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
$CMD > gpurun_out/r02_c8_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c8_launches.csv $CMD > gpurun_out/r02_c8_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_msm_affine -s 2 -c 1 -o gpurun_out/prof_msm_affine_r02_fused $CMD > gpurun_out/r02_c8_ncu.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -2
