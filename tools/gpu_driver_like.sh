#!/usr/bin/env bash
# What the driver runs at round end, in its order: GPU tests, smoke(), the bench (both arms).  usage: gpu_driver_like.sh <tag>
T=${1:-drv}; O=gpurun_out; mkdir -p $O
( time python -m pytest tests/ -x -q -m gpu ) > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; grep -E "smoke|real" $O/${T}_smoke.log | cut -c1-200
( time python bench.py --impl reference ) > $O/${T}_bench_ref.log 2> $O/${T}_bench_ref.err; echo "ref rc=$?"; tail -c 400 $O/${T}_bench_ref.log; grep real $O/${T}_bench_ref.err
( time python bench.py ) > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench rc=$?"; python tools/show_bench.py $O/${T}_bench.log; grep real $O/${T}_bench.err
