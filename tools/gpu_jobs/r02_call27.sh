#!/bin/bash
# Round 2, GPU call 27 (1 x B200): r-power terms with the scalars split into 128-bit halves (2^128 P precomputed beside
# the transcript hash): verify suite, callers, C++ host, stage trace.
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_verify.py tests/test_gpu_callers.py tests/test_gpu_cpp_host.py -m gpu -x -q ) > gpurun_out/r02_c27_pytest_verify.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c27_pytest_verify.txt; tail -4 gpurun_out/r02_c27_pytest_verify.txt
RAIKO_KZG_VERIFY_TRACE=1 python tests/tools/verify_trace.py 4096 > gpurun_out/r02_c27_verify.txt 2>&1; tail -7 gpurun_out/r02_c27_verify.txt
