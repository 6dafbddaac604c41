#!/usr/bin/env bash
# One gpurun call: GPU parity tests, the bench line (both arms), and the ncu launch list of the same step.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.log
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
