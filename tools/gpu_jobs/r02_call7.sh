#!/bin/bash
# Round 2, GPU call 7 (1 x B200): GPU suite with the fused limb passes (default now) and the FMA-pipe SHA-256
# additions; IMAD operand-form microbenchmark; configs (6-blob latency, 4096 commit, verify); ncu --set full of
# k_fr_eval_quot and k_sha_blob_duo.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/r02_c7_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c7_pytest.txt; tail -4 gpurun_out/r02_c7_pytest.txt
tools/ubench/wide_operands3 > gpurun_out/r02_c7_wide_operands3.txt 2>&1; cat gpurun_out/r02_c7_wide_operands3.txt
python bench.py --batch 9472 --steps 2 --warmup 2 --cpu-sample 16 > gpurun_out/r02_c7_bench_9472.json 2> gpurun_out/r02_c7_bench_9472.err
echo "bench rc=$?"; python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_c7_bench_9472.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["kernel_ms"]); print(json.dumps(d.get("configs"))[:1500])
P
RAIKO_KZG_VERIFY_TRACE=1 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c7_verify_lanes.txt 2>&1; tail -7 gpurun_out/r02_c7_verify_lanes.txt
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
ncu --set full --clock-control none --import-source on -k regex:k_fr_eval_quot -s 1 -c 1 -o gpurun_out/prof_fr_r02 $CMD > gpurun_out/r02_c7_ncu_fr.log 2>&1; echo "ncu fr rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_sha_blob_duo -s 1 -c 1 -o gpurun_out/prof_sha_r02 $CMD > gpurun_out/r02_c7_ncu_sha.log 2>&1; echo "ncu sha rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
