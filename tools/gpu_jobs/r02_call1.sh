#!/bin/bash
# Round 2, GPU call 1 (1 x B200): full GPU suite, the bench line, window-width sweep, launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_c1_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c1_pytest.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_c1_pytest.txt
tail -5 gpurun_out/r02_c1_pytest.txt
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_c1_bench.err
for wb in 13 14; do
  python bench.py --batch 9472 --window-bits $wb --steps 2 --warmup 1 --no-e2e --no-configs --cpu-sample 16 > gpurun_out/r02_c1_bench_c$wb.json 2> gpurun_out/r02_c1_bench_c$wb.err
  echo "bench c=$wb rc=$?"
done
RAIKO_KZG_WAVE_TAIL=0 python bench.py --batch 8192 --steps 2 --warmup 1 --no-configs --cpu-sample 16 > gpurun_out/r02_c1_bench_8192_notail.json 2> gpurun_out/r02_c1_bench_8192_notail.err
python bench.py --batch 8192 --steps 2 --warmup 1 --no-configs --cpu-sample 16 > gpurun_out/r02_c1_bench_8192_tail.json 2> gpurun_out/r02_c1_bench_8192_tail.err
echo "8192 rc=$?"
CMD="python bench.py --batch 4736 --steps 1 --warmup 1 --no-e2e --no-configs --cpu-sample 16"
$CMD > gpurun_out/r02_c1_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_c1_launches.csv $CMD > gpurun_out/r02_c1_ncu.log 2>&1
echo "ncu rc=$?"
