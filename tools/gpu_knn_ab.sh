#!/usr/bin/env bash
# A/B of the kNN kernel families under the bench workloads. usage: gpu_knn_ab.sh <tag> [algos] [workloads]
T=${1:-knn}; ALGOS=${2:-"1 0 2"}; WL=${3:-"train garden eval"}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_knn.py -m gpu -q -rA > $O/${T}_pytest.log 2>&1; grep -E "passed|failed|ms \(points" $O/${T}_pytest.log | cut -c1-300
for a in $ALGOS; do
  for w in $WL; do
    extra=""; [ $w != train ] && extra="--workload $w"
    SPF_KNN_ALGO=$a timeout 600 python bench.py $extra --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/${T}_a${a}_$w.log 2> $O/${T}_a${a}_$w.err
    python - <<PY
import json
for line in open("$O/${T}_a${a}_$w.log"):
    if line.startswith("{"):
        d=json.loads(line); k=d.get("kernels_ms_per_step",{})
        print("algo $a $w ms/step", round(d["ms_per_step"],3), {n: round(v,3) for n,v in k.items() if "knn" in n})
PY
  done
done
