#!/bin/bash
# Round 2, GPU call 3 (1 x B200): suite after the verification rewrite, verify stage traces, e2e trace after the
# hash-placement fix, Fr kernel timing.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c3_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c3_pytest.txt
tail -4 gpurun_out/r02_c3_pytest.txt
RAIKO_KZG_VERIFY_TRACE=1 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c3_verify_lanes.txt 2>&1
RAIKO_KZG_VERIFY_TRACE=1 RAIKO_KZG_PAIRING_LANES=0 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c3_verify_scalar.txt 2>&1
tail -8 gpurun_out/r02_c3_verify_lanes.txt; tail -3 gpurun_out/r02_c3_verify_scalar.txt
RAIKO_KZG_TRACE=1 python tests/tools/e2e_trace.py 8192 3 > gpurun_out/r02_c3_trace_8192.txt 2>&1
tail -5 gpurun_out/r02_c3_trace_8192.txt
python bench.py --batch 9472 --steps 2 --warmup 2 --cpu-sample 16 > gpurun_out/r02_c3_bench_9472.json 2> gpurun_out/r02_c3_bench_9472.err
echo "bench rc=$?"
./tools/ubench/fpmul_dfma > gpurun_out/r02_c3_fpmul_dfma.txt 2>&1
tail -8 gpurun_out/r02_c3_fpmul_dfma.txt
