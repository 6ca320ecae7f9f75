#!/bin/bash
# Round 2, GPU call 2 (2 x B200): whole GPU suite with the multi-GPU tests live, strong-scaling bench at N=2
# (incl. the single-process two-device leg), e2e chunk timelines.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02_c2_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/r02_c2_pytest_2gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c2_pytest_2gpu.txt
tail -4 gpurun_out/r02_c2_pytest_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_c2_bench_n2.json 2> gpurun_out/r02_c2_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/r02_c2_bench_n2.err
for tail in 1 0; do
  RAIKO_KZG_TRACE=1 RAIKO_KZG_WAVE_TAIL=$tail python tests/tools/e2e_trace.py 8192 3 > gpurun_out/r02_c2_trace_8192_tail$tail.txt 2>&1
done
RAIKO_KZG_TRACE=1 python tests/tools/e2e_trace.py 32768 2 > gpurun_out/r02_c2_trace_32768.txt 2>&1
tail -3 gpurun_out/r02_c2_trace_8192_tail1.txt
