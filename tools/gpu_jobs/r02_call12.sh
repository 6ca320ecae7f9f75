#!/bin/bash
# Round 2, GPU call 12 (1 x B200): chunk-size sweep (the blob hash is one latency-bound chain per blob: 2.2 ms per chunk
# whatever its size), device-resident and host-buffer legs.
set -u
mkdir -p gpurun_out
for chunk in 4736 9472 14208; do
  RAIKO_KZG_CHUNK=$chunk python bench.py --batch 28416 --steps 2 --warmup 2 --no-configs --cpu-sample 16 > gpurun_out/r02_c12_bench_chunk$chunk.json 2> gpurun_out/r02_c12_bench_chunk$chunk.err
  python - <<P
import json
d = json.loads(open("gpurun_out/r02_c12_bench_chunk$chunk.json").read().strip().splitlines()[-1])
print("chunk $chunk: value %.0f e2e %.0f kernel_ms %s launches %d" % (d["value"], d["e2e"]["value"], {k: round(v, 1) for k, v in d["kernel_ms"].items()}, d["gpu_launches"]))
P
done 2>&1 | tee gpurun_out/r02_c12_chunk_sweep.txt
