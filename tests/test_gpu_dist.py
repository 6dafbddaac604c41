"""GPU, multi-device (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`,
log committed under profiles/): the ray-sharded NCCL training step equals the one-big-batch step."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sharded_nccl_step_equals_big_batch_step(precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (NCCL)")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py"), precision]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, (res.returncode, res.stdout[-2000:], res.stderr[-3000:])
    d = json.loads(lines[-1])
    print("sharded vs big batch:", json.dumps(d))
    assert d["ranks_identical"]
    assert len(set(d["hit_rays_per_rank"])) > 1, d["hit_rays_per_rank"]     # the shards really see different hit counts
    tol = 1e-5 if precision == "fp32" else 1e-3   # same kernels on both sides: only the summation order / tile split differs
    for k, (a, b) in d["terms"].items():
        assert abs(a - b) <= tol * max(1.0, abs(b)), (k, a, b)
    assert d["grad_rel_err"] < (2e-5 if precision == "fp32" else 2e-3), d["grad_rel_err"]
    # Adam's first step moves every entry by ~lr * g / (|g| + eps): entries with |g| ~ eps are ill-conditioned, so bound
    # the mean difference tightly and the max by one step size
    assert d["adam_delta_mean_abs_diff"] <= 2e-3 * d["adam_delta_mean_abs"], d
    assert d["adam_delta_max_abs_diff"] <= 1.01 * d["adam_delta_absmax"], d
    # the overlapped exchange (early all-reduce of the colour latents inside the backward), eager and graph-replayed,
    # takes the same step as the single all-reduce after the backward
    for k in ("overlap_vs_serial_mean_abs_diff", "graph_vs_eager_mean_abs_diff", "overlap_vs_manual_mean_abs_diff"):
        assert d[k] <= 2e-3 * d["adam_delta_mean_abs"], (k, d)
