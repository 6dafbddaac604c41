"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port timed on the host cores) prints ONE JSON line
with the keys the driver's contract names; under torchrun every rank but 0 exits without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-rays", "8"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train rays/s (fwd+bwd)" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_no_work():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_other_rooflines_from_a_fake_profile():
    """bench.other_rooflines (the per-kernel algorithmic-work rooflines beside the headline one) on a synthetic profile:
    every kernel it knows gets achieved / peak / frac, missing kernels are skipped, nothing needs a GPU."""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    names = ["spf_color_fwd_tc", "spf_color_bwd_tc", "spf_wgrad_tc_multi", "spf_head_fwd_tc", "spf_head_bwd_tc", "spf_knn_slots",
             "spf_composite_fwd", "spf_adam_step", "spf_grad_sumsq", "spf_sdf_bwd", "spf_sampler_iter", "spf_tv_fwd_bwd"]
    prof = {k: {"ms": 2.0, "calls": 2, "max_ms": 1.0} for k in names}

    class Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1000))
            self.neural_pts = torch.zeros(100, 3)
            self._last = {"t": torch.zeros(4096, 80)}
            self._bench_cq = {"queries": 1000, "candidates": 500000}

    pk = bench.peaks()
    out = bench.other_rooflines(prof, 2, 1.2e6, Fake(), pk)
    assert set(out) == set(names)
    for k, v in out.items():
        assert v["unit"] in ("GB/s", "TFLOP/s") and v["achieved"] > 0 and abs(v["ms_per_step"] - 1.0) < 1e-9
        if "resident" in v:     # L2-resident kernels carry no fraction of the HBM peak
            assert v["peak"] is None and v["frac"] is None and k in ("spf_knn_slots", "spf_sdf_bwd")
        else:
            assert v["peak"] > 0 and abs(v["frac"] - v["achieved"] / v["peak"]) < 1e-12
    del prof["spf_tv_fwd_bwd"]
    assert "spf_tv_fwd_bwd" not in bench.other_rooflines(prof, 2, 1.2e6, Fake(), pk)
