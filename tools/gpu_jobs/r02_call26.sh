#!/bin/bash
# Round 2, GPU call 26 (1 x B200): run-to-run stability of the default bench line (three separate processes).
set -u
mkdir -p gpurun_out
for i in 1 2 3; do
  python bench.py --steps 3 --warmup 3 --cpu-sample 16 > gpurun_out/r02_c26_bench_run$i.json 2> gpurun_out/r02_c26_bench_run$i.err
  python - <<P
import json
d = json.loads(open("gpurun_out/r02_c26_bench_run$i.json").read().strip().splitlines()[-1])
c = d["configs"]; r = d["roofline"]
print("run $i: value %.0f e2e %.0f | frac %.3f slot_frac %.3f pipe_busy %.3f | 6 blobs %.2f ms, commit-only %.0f/s, verify %.1f ms | clocks %s" % (
    d["value"], d["e2e"]["value"], r["frac"], r["slot_frac"], r["pipe_busy"], c["config1_6_blobs_commit_prove_ms"]["best"],
    c["config2_4096_blobs_commit_only"]["blobs_per_s_best"], 1e3 * c["config4_4096_blobs_verify_batch"]["seconds_best"], d["clocks"]))
P
done 2>&1 | tee gpurun_out/r02_c26_stability.txt
