"""Golden vectors for the neural-point ingestion from the REFERENCE's own functions
``spurfies/model/utils.py::construct_vox_points_closest`` / ``voxelize`` (:6-59), imported from /root/reference.

The reference needs ``torch_scatter`` (absent in this image).  It is replaced by a ~20-line pure-torch shim with
torch_scatter's documented CPU semantics: ``scatter_mean`` = fp32 scatter-add in index order / count,
``scatter_min`` = (min value, index of the FIRST element attaining it).  ``.cuda()`` is redirected to CPU as in
make_golden.py.  Runs only in the authoring container; the tests use the committed tests/golden/ingest.pt.

Usage:  python tests/golden/make_golden_ingest.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("SPF_REFERENCE_ROOT", "/root/reference")


def scatter_mean(src, index, dim=0):
    assert dim == 0
    n = int(index.max()) + 1
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, src)
    cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones(len(index), dtype=src.dtype))
    return out / cnt.clamp(min=1).reshape((n,) + (1,) * (src.dim() - 1))


def scatter_min(src, index, dim=0):
    assert dim == 0 and src.dim() == 1
    n = int(index.max()) + 1
    best = torch.full((n,), float("inf"), dtype=src.dtype).scatter_reduce_(0, index, src, reduce="amin")
    m = len(src)
    cand = torch.where(src == best[index], torch.arange(m), torch.full((m,), m))
    arg = torch.full((n,), m, dtype=torch.long).scatter_reduce_(0, index, cand, reduce="amin")
    return best, arg


def main():
    ts = types.ModuleType("torch_scatter")
    ts.scatter_mean, ts.scatter_min = scatter_mean, scatter_min
    sys.modules["torch_scatter"] = ts
    sys.modules.setdefault("plyfile", types.ModuleType("plyfile"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    _zeros = torch.zeros
    torch.zeros = lambda *a, **k: _zeros(*a, **{kk: ("cpu" if kk == "device" else v) for kk, v in k.items()})
    sys.path.insert(0, REF)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_model_utils", os.path.join(REF, "spurfies", "model", "utils.py"))
    U = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(U)
    from spurfies_b200 import scenes
    gold = {"cases": []}
    for n, seed, vox_res in ((20000, 3, 50), (30000, 5, 60), (1000, 3, 7)):
        pts = scenes.dtu_like(n, seed=seed, radii=(0.35, 0.5))["pts"]
        cen, gidx, midx = U.construct_vox_points_closest(pts.clone(), vox_res)
        kept, idx = U.voxelize(pts.clone(), vox_res)
        assert torch.equal(idx, midx) and torch.equal(kept, pts[idx])
        gold["cases"].append({"n": n, "seed": seed, "radii": (0.35, 0.5), "vox_res": vox_res,
                              "pts_checksum": float(pts.double().abs().sum()), "centroid": cen.clone(),
                              "grid_idx": gidx.clone(), "min_idx": midx.clone()})
        print(n, vox_res, "->", len(midx), "voxels")
    path = os.path.join(ROOT, "tests", "golden", "ingest.pt")
    torch.save(gold, path)
    print(f"wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
