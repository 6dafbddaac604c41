#!/usr/bin/env bash
# quick GPU check: parity tests (with a timeout in case a kernel hangs) + a short bench. Usage: gpu_quick.sh tag [pytest -k expr]
T=${1:-q}; K=${2:-}
O=gpurun_out; mkdir -p $O
if [ -n "$K" ]; then
  timeout 600 python -m pytest tests -m gpu -q -rA -x -k "$K" > $O/${T}_pytest.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -q -rA > $O/${T}_pytest.log 2>&1
fi
echo "pytest rc=$?" >> $O/${T}_pytest.log
grep -E "passed|failed|FAILED|ERROR|rc=|tc vs|emulation|bf16 mode" $O/${T}_pytest.log | cut -c1-600 | tail -30
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/${T}_bench.log 2> $O/${T}_bench.err
  python - <<PY
import json
for line in open("$O/${T}_bench.log"):
    if line.startswith("{"):
        d=json.loads(line); print("value",d["value"],"ms",d["ms_per_step"],"e2e ms",d["e2e"]["ms_per_step"],"graph",d["config"]["cuda_graph"], d["config"]["cuda_graph_note"]); print({k:round(v,3) for k,v in d["kernels_ms_per_step"].items()}); print(d["roofline"])
PY
  tail -3 $O/${T}_bench.err
fi
