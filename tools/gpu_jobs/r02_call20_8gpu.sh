#!/bin/bash
# Round 2, GPU call 20 (8 x B200, final library): BASELINE.json configs[3] as written -- ONE 65,536-blob batch sharded over 8 and 4
# GPUs (strong scaling, per-rank leg + the single-process multi-device leg), and the library's own sharding test
# on all 8 devices with a batch size no device count divides.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02_c20_gpu.txt 2>&1
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -rs -k "multi_gpu" ) > gpurun_out/r02_c20_pytest_8gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c20_pytest_8gpu.txt
tail -4 gpurun_out/r02_c20_pytest_8gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02_c20_bench_n8.json 2> gpurun_out/r02_c20_bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/r02_c20_bench_n8.err; cut -c1-400 gpurun_out/r02_c20_bench_n8.json
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02_c20_bench_n4.json 2> gpurun_out/r02_c20_bench_n4.err
echo "bench n4 rc=$?"; tail -3 gpurun_out/r02_c20_bench_n4.err; cut -c1-400 gpurun_out/r02_c20_bench_n4.json
CUDA_VISIBLE_DEVICES=0,1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_c20_bench_n2.json 2> gpurun_out/r02_c20_bench_n2.err
echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/r02_c20_bench_n2.json
