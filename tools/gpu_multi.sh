#!/usr/bin/env bash
# multi-GPU call: NCCL equivalence test + weak-scaling bench (fp32 and bf16 gradient exchange) + garden strong scaling
# usage: bash tools/gpu_multi.sh <tag> <ngpus> [steps...]   steps: test bench bench16 garden garden16 eval mesh
T=${1:-m}; N=${2:-2}; shift 2 || true
STEPS=${*:-test bench bench16 garden}
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
for s in $STEPS; do
  case $s in
    test) timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -rA > $O/${T}_dist_pytest.log 2>&1; grep -E "passed|failed|skipped|sharded vs" $O/${T}_dist_pytest.log | cut -c1-1500 ;;
    bench) timeout 600 $RUN --steps 20 --warmup 3 > $O/${T}_bench_n$N.log 2> $O/${T}_bench_n$N.err; python tools/show_bench.py $O/${T}_bench_n$N.log ;;
    overlap) timeout 600 $RUN --steps 20 --warmup 3 --dp-overlap > $O/${T}_overlap_n$N.log 2> $O/${T}_overlap_n$N.err; python tools/show_bench.py $O/${T}_overlap_n$N.log ;;
    noexch) SPF_DP_DIAG_SKIP_REDUCE=1 timeout 600 $RUN --steps 20 --warmup 3 > $O/${T}_noexch_n$N.log 2> $O/${T}_noexch_n$N.err; python tools/show_bench.py $O/${T}_noexch_n$N.log ;;
    nvls) NCCL_ALGO=NVLS NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 600 $RUN --steps 20 --warmup 3 > $O/${T}_nvls_n$N.log 2> $O/${T}_nvls_n$N.err; python tools/show_bench.py $O/${T}_nvls_n$N.log; grep -i "nvls\|algo" $O/${T}_nvls_n$N.log $O/${T}_nvls_n$N.err | head -8 | cut -c1-200 ;;
    bench16) timeout 600 $RUN --steps 20 --warmup 3 --grad-compress bf16 > $O/${T}_bench16_n$N.log 2> $O/${T}_bench16_n$N.err; python tools/show_bench.py $O/${T}_bench16_n$N.log ;;
    garden) timeout 900 $RUN --workload garden --steps 10 --warmup 3 > $O/${T}_garden_n$N.log 2> $O/${T}_garden_n$N.err; python tools/show_bench.py $O/${T}_garden_n$N.log ;;
    garden16) timeout 900 $RUN --workload garden --steps 10 --warmup 3 --grad-compress bf16 > $O/${T}_garden16_n$N.log 2> $O/${T}_garden16_n$N.err; python tools/show_bench.py $O/${T}_garden16_n$N.log ;;
    eval) timeout 900 $RUN --workload eval --steps 3 --warmup 3 > $O/${T}_eval_n$N.log 2> $O/${T}_eval_n$N.err; python tools/show_bench.py $O/${T}_eval_n$N.log ;;
    mesh) timeout 900 $RUN --workload mesh --steps 2 --warmup 3 > $O/${T}_mesh_n$N.log 2> $O/${T}_mesh_n$N.err; python tools/show_bench.py $O/${T}_mesh_n$N.log ;;
  esac
done
