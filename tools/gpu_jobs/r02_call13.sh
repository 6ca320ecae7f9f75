#!/bin/bash
# Round 2, GPU call 13 (1 x B200): lazy sums / fold in the CTA-wide pairing check; chunk-size sweep.
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_verify.py tests/test_gpu_cpp_host.py -m gpu -x -q ) > gpurun_out/r02_c13_pytest_verify.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c13_pytest_verify.txt; tail -4 gpurun_out/r02_c13_pytest_verify.txt
for lanes in 2 1; do
  RAIKO_KZG_PAIRING_LANES=$lanes RAIKO_KZG_VERIFY_TRACE=1 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c13_verify_lanes$lanes.txt 2>&1; tail -3 gpurun_out/r02_c13_verify_lanes$lanes.txt
done
bash tools/gpu_jobs/r02_call12.sh
