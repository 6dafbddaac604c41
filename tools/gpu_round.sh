#!/usr/bin/env bash
# One gpurun call: GPU parity tests, the bench line (both arms), the ncu launch list of the same step and one
# `--set full` capture of the dominant kernel.  Usage: bash tools/gpu_round.sh [tag]   (outputs under gpurun_out/<tag>_*)
set -u
T=${1:-run}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/${T}_gpu.txt 2>&1
( time python -m pytest tests -m gpu -q -rA ) > $O/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest_gpu.log
grep -E "passed|failed|PASSED|FAILED|rc=" $O/${T}_pytest_gpu.log | tail -45
( time python bench.py --steps 10 --warmup 3 ) > $O/${T}_bench.log 2> $O/${T}_bench.err
tail -c 2500 $O/${T}_bench.log
if [ "${SKIP_REF:-0}" != "1" ]; then
  ( time python bench.py --impl reference --steps 2 --warmup 1 ) > $O/${T}_bench_ref.log 2> $O/${T}_bench_ref.err
  tail -c 600 $O/${T}_bench_ref.log
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-reference-gpu > $O/${T}_bench_ncu.log 2>&1
  echo "ncu launches rc=$?"; wc -l $O/${T}_launches.csv
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNEL:-k_sdf_tc2|k_knn_slots|k_knn_points|k_color_fwd_tc2|k_color_bwd_tc2|k_wgrad|k_composite_fwd|k_sampler_iter|k_head_fwd_tc2|k_head_bwd_tc2|k_sdf_bwd}" -s ${NCU_SKIP:-66} -c ${NCU_COUNT:-22} -f -o $O/${T}_top \
    python bench.py --steps 2 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-reference-gpu > $O/${T}_bench_ncu_full.log 2>&1
  echo "ncu full rc=$?"; ls -la $O/${T}_top.ncu-rep
fi
