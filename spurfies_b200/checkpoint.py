"""Checkpoint files around the hot path, in the reference's formats (spurfies/train.py), so that a run can move
between the two implementations in either direction.  Host-side only (torch.save / torch.load of state dicts).

* the local-prior file ``ckpt/local_prior.pt`` -> the frozen ``F_geometry`` / ``T`` weights: the key mapping of
  train.py:123-140;
* ``<exp>/checkpoints/ModelParameters/{<epoch>,latest}.pth`` = ``{"epoch", "model_state_dict", "iter_step"}`` and
  ``<exp>/checkpoints/OptimizerParameters/{<epoch>,latest}.pth`` = ``{"epoch", "optimizer_state_dict"}``
  (train.py:292-328 ``save_checkpoints``, :222-241 ``load_from_dir``).  ``PointVolSDF`` has the reference's parameter /
  buffer names and shapes (tests/test_checkpoint.py checks them against the reference class), ``FusedAdam.state_dict``
  has ``torch.optim.Adam``'s layout, so both files load with the other side's ``load_state_dict``.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

MODEL_SUBDIR, OPTIM_SUBDIR = "ModelParameters", "OptimizerParameters"   # train.py:87-93


def prior_to_model_state(prior: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """train.py:125-139.  ``prior`` = the ``model_state_dict`` of ``ckpt/local_prior.pt``.  After dropping
    ``sdf_features`` the i-th entry whose key contains ``local_sdf_field`` becomes
    ``F_geometry.{0,0,2,2,4,4,6,6,8,8}[i].<key components from the fifth on>`` (weight / bias of the five linears, in file
    order) and ``density_branch.{weight,bias}`` become ``T.0.{weight,bias}``.  The index i counts EVERY remaining entry,
    as the reference's ``enumerate(prior.items())`` does."""
    prior = {k: v for k, v in prior.items() if k != "sdf_features"}
    layer = [0, 0, 2, 2, 4, 4, 6, 6, 8, 8]
    out: Dict[str, torch.Tensor] = {}
    for i, (k, v) in enumerate(prior.items()):
        if "local_sdf_field" in k:
            if i >= len(layer):
                raise ValueError(f"local prior: entry {i} ('{k}') is beyond the five linears of F_geometry")
            out[f"F_geometry.{layer[i]}." + ".".join(k.split(".")[4:])] = v
        if "density_branch.weight" in k:
            out["T.0.weight"] = v
        if "density_branch.bias" in k:
            out["T.0.bias"] = v
    return out


def load_prior(model: torch.nn.Module, prior, freeze: bool = True) -> Dict[str, torch.Tensor]:
    """Load the local prior (a path to ``local_prior.pt`` or its already loaded dict) into ``model`` and, as
    train.py:148-154 does, switch the gradients of ``F_geometry`` / ``T`` off.  Returns the mapped entries.  Unlike the
    reference's ``strict=False`` call, an entry whose name or shape does not fit the model raises."""
    if isinstance(prior, (str, os.PathLike)):
        prior = torch.load(prior, map_location="cpu", weights_only=True)
    sd = prior_to_model_state(prior["model_state_dict"] if "model_state_dict" in prior else prior)
    own = model.state_dict()
    for k, v in sd.items():
        if k not in own or tuple(own[k].shape) != tuple(v.shape):
            raise ValueError(f"local prior: '{k}' {tuple(v.shape)} does not fit the model "
                             f"({tuple(own[k].shape) if k in own else 'no such entry'})")
    model.load_state_dict(sd, strict=False)
    if freeze:
        for name, p in model.named_parameters():
            if "F_geometry" in name or "T.0" in name:
                p.requires_grad_(False)
    return sd


def reference_optimizer_layout(opt_sd: Dict) -> Dict:
    """The reference builds its Adam with TWO parameter groups (train.py:168-189): an empty one (``sdf_feat``, lr 1e-2)
    and the trainable tensors; ``torch.optim.Adam.load_state_dict`` insists on the same number and sizes of groups.  A
    single-group state dict (``FusedAdam.state_dict()``) gets the empty group prepended; anything else is returned as is."""
    groups = opt_sd["param_groups"]
    if len(groups) != 1:
        return opt_sd
    empty = dict(groups[0])
    empty["params"], empty["lr"] = [], 1e-2
    return {"state": opt_sd["state"], "param_groups": [empty, groups[0]]}


def save_checkpoints(checkpoints_path: str, epoch: int, model: torch.nn.Module, optimizer, iter_step: int,
                     latest_only: bool = False, reference_layout: bool = True) -> None:
    """train.py:292-328: ``latest.pth`` always, ``<epoch>.pth`` unless ``latest_only``.  ``optimizer`` is anything with
    a ``state_dict()`` (``FusedAdam``, or ``TrainStep.opt``); with ``reference_layout`` its file carries the reference's
    two parameter groups, so the reference's own ``load_from_dir`` accepts it (``FusedAdam.load_state_dict`` reads either
    layout)."""
    names = ["latest"] if latest_only else ["latest", str(epoch)]
    for sub in (MODEL_SUBDIR, OPTIM_SUBDIR):
        os.makedirs(os.path.join(checkpoints_path, sub), exist_ok=True)
    for n in names:
        torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "iter_step": int(iter_step)},
                   os.path.join(checkpoints_path, MODEL_SUBDIR, n + ".pth"))
        osd = optimizer.state_dict()
        torch.save({"epoch": epoch, "optimizer_state_dict": reference_optimizer_layout(osd) if reference_layout else osd},
                   os.path.join(checkpoints_path, OPTIM_SUBDIR, n + ".pth"))


def load_from_dir(checkpoints_path: str, model: torch.nn.Module, optimizer=None, checkpoint: str = "latest",
                  map_location: Optional[str] = None) -> Dict[str, int]:
    """train.py:222-241: strict load of the model file, then (if given) the optimiser file.  Returns
    ``{"epoch", "iter_step"}``; a trainer resumes its schedule from ``iter_step`` (``TrainStep.iter_step``)."""
    m = torch.load(os.path.join(checkpoints_path, MODEL_SUBDIR, str(checkpoint) + ".pth"), map_location=map_location,
                   weights_only=False)
    model.load_state_dict(m["model_state_dict"])
    if optimizer is not None:
        o = torch.load(os.path.join(checkpoints_path, OPTIM_SUBDIR, str(checkpoint) + ".pth"), map_location=map_location,
                       weights_only=False)
        optimizer.load_state_dict(o["optimizer_state_dict"])
    return {"epoch": int(m["epoch"]), "iter_step": int(m.get("iter_step", 0))}
