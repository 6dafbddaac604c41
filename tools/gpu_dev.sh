#!/usr/bin/env bash
# development: tensor-core parity tests + short bench + colour forward timeline.  usage: gpu_dev.sh <tag> [fwd|bwd|none]
T=${1:-dev}; W=${2:-fwd}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rA -x -k "tc or hotpath or reference_path or dist" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=" $O/${T}_pytest.log | cut -c1-300 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/${T}_bench.log 2> $O/${T}_bench.err; python tools/show_bench.py $O/${T}_bench.log
if [ "$W" != "none" ]; then SPF_LIBRARY=spurfies_b200/csrc/libspurfies_b200_tl.so timeout 300 python tools/timeline_color.py $W > $O/${T}_tl_$W.log 2>&1; head -16 $O/${T}_tl_$W.log; fi
