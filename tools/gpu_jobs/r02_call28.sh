#!/bin/bash
# Round 2, GPU call 28 (1 x B200): final state of the library: whole GPU suite, smoke(), default bench line, reference arm.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/r02_c28_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c28_pytest.txt; tail -4 gpurun_out/r02_c28_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c28_smoke.txt 2>&1; tail -1 gpurun_out/r02_c28_smoke.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c28_bench.json 2> gpurun_out/r02_c28_bench.err
echo "bench rc=$?"; python - <<'P'
import json
d = json.loads(open("gpurun_out/r02_c28_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity_checked_blobs"]); c = d["configs"]; r = d["roofline"]
print(c["config1_6_blobs_commit_prove_ms"]["best"], c["config2_4096_blobs_commit_only"]["blobs_per_s_best"], c["config4_4096_blobs_verify_batch"]["seconds_best"], r["frac"], r["slot_frac"], r["pipe_busy"], r["model_frac"])
P
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c28_bench_ref.json 2> gpurun_out/r02_c28_bench_ref.err; cut -c1-160 gpurun_out/r02_c28_bench_ref.json
