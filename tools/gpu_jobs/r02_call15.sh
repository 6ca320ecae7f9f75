#!/bin/bash
# Round 2, GPU call 15 (1 x B200): whole GPU suite and the default bench line with the Karatsuba product rows
# (RK_MUL_FORM=3) as the library default.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/r02_c15_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c15_pytest.txt; tail -4 gpurun_out/r02_c15_pytest.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_c15_bench.json 2> gpurun_out/r02_c15_bench.err
echo "bench rc=$?"; python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_c15_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "kernel_ms", d["kernel_ms"], "parity", d["parity_checked_blobs"]); c = d["configs"]
print(c["config1_6_blobs_commit_prove_ms"]["best"], c["config2_4096_blobs_commit_only"]["blobs_per_s_best"], c["config4_4096_blobs_verify_batch"]["seconds_best"])
P
RAIKO_KZG_VERIFY_TRACE=1 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c15_verify.txt 2>&1; tail -7 gpurun_out/r02_c15_verify.txt
