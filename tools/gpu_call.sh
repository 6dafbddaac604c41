#!/usr/bin/env bash
# One gpurun call for development: parity tests, headline bench, the other BASELINE workloads, colour-kernel timelines.
# usage: bash tools/gpu_call.sh <tag> [steps...]   steps: tests bench eval mesh garden tl ref
set -u
T=${1:-dev}; shift || true
STEPS=${*:-tests bench eval mesh garden tl}
O=gpurun_out; mkdir -p $O
for s in $STEPS; do
  case $s in
    tests) timeout 900 python -m pytest tests -m gpu -q -rA > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
           grep -E "passed|failed|FAILED|ERROR|rc=" $O/${T}_pytest.log | cut -c1-300 | tail -15 ;;
    bench) timeout 600 python bench.py --steps 10 --warmup 3 > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -c 1500 $O/${T}_bench.log; tail -3 $O/${T}_bench.err ;;
    ref)   timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.log 2> $O/${T}_bench_ref.err; tail -c 300 $O/${T}_bench_ref.log ;;
    eval)  timeout 600 python bench.py --workload eval --steps 3 --warmup 3 > $O/${T}_bench_eval.log 2> $O/${T}_bench_eval.err; echo "eval rc=$?"; tail -c 1200 $O/${T}_bench_eval.log; tail -3 $O/${T}_bench_eval.err ;;
    mesh)  timeout 600 python bench.py --workload mesh --steps 2 --warmup 3 > $O/${T}_bench_mesh.log 2> $O/${T}_bench_mesh.err; echo "mesh rc=$?"; tail -c 1200 $O/${T}_bench_mesh.log; tail -3 $O/${T}_bench_mesh.err ;;
    garden) timeout 600 python bench.py --workload garden --steps 5 --warmup 3 > $O/${T}_bench_garden.log 2> $O/${T}_bench_garden.err; echo "garden rc=$?"; tail -c 1500 $O/${T}_bench_garden.log; tail -3 $O/${T}_bench_garden.err ;;
    tl)    for w in fwd bwd; do SPF_LIBRARY=spurfies_b200/csrc/libspurfies_b200_tl.so timeout 300 python tools/timeline_color.py $w > $O/${T}_tl_$w.log 2>&1; echo "tl $w rc=$?"; head -30 $O/${T}_tl_$w.log; done ;;
  esac
done
