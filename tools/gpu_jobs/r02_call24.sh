#!/bin/bash
# Round 2, GPU call 24 (1 x B200): the new pairing-kernel agreement / determinism test and the verify suite.
set -u
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_verify.py -m gpu -x -q ) > gpurun_out/r02_c24_pytest_verify.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c24_pytest_verify.txt; tail -5 gpurun_out/r02_c24_pytest_verify.txt
