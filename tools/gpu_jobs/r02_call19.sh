#!/bin/bash
# Round 2, GPU call 19 (1 x B200): the library as it will ship (RK_MUL_FORM=4, four lockstep groups, fused passes):
# whole GPU suite, smoke(), the default bench line, ncu launch list, ncu --set full of k_msm_affine / k_fr_eval_quot / k_msm.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/r02_c19_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c19_pytest.txt; tail -4 gpurun_out/r02_c19_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c19_smoke.txt 2>&1; tail -1 gpurun_out/r02_c19_smoke.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c19_bench.json 2> gpurun_out/r02_c19_bench.err
echo "bench rc=$?"; python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_c19_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["kernel_ms"], "parity", d["parity_checked_blobs"]); c = d["configs"]
print(c["config1_6_blobs_commit_prove_ms"]["best"], c["config2_4096_blobs_commit_only"]["blobs_per_s_best"], c["config4_4096_blobs_verify_batch"]["seconds_best"])
P
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c19_bench_ref.json 2> gpurun_out/r02_c19_bench_ref.err; cut -c1-200 gpurun_out/r02_c19_bench_ref.json
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c19_launches.csv $CMD > gpurun_out/r02_c19_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_msm_affine -s 2 -c 1 -o gpurun_out/prof_msm_affine_r02_ship $CMD > gpurun_out/r02_c19_ncu_aff.log 2>&1; echo "ncu affine rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_fr_eval_quot -s 1 -c 1 -o gpurun_out/prof_fr_r02_ship $CMD > gpurun_out/r02_c19_ncu_fr.log 2>&1; echo "ncu fr rc=$?"
CMDX="python bench.py --batch 2368 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
RAIKO_KZG_MSM_AFFINE=0 ncu --set full --clock-control none --import-source on -k regex:k_msm$ -s 2 -c 1 -o gpurun_out/prof_msm_xyzz_r02_ship $CMDX > gpurun_out/r02_c19_ncu_xyzz.log 2>&1; echo "ncu xyzz rc=$?"
