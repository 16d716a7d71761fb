"""Checkpoint / resume (reference ``src/sync.jl:156-161,166``).

The reference saves ``cpu(model)`` to ``weights/<p>/resnet_50_cycle_<n>_<time>.bson``
every 20 cycles (model only, directory must pre-exist, quirk Q11) and resumes
in-memory by passing ``sts=`` back in.  Here a checkpoint is one atomic file with
{model state_dict, optimiser state tree, step, RNG state, shard permutation,
user metadata}; replicas are identical so rank/replica 0 writes.
"""
from __future__ import annotations

import os
import tempfile
import time
from typing import Any, Dict, Optional

import torch

from ..tree import fmap


def _cpu(tree):
    return fmap(lambda t: t.detach().cpu() if isinstance(t, torch.Tensor) else t, tree)


def save_checkpoint(path: str, model: torch.nn.Module, opt_state=None, step: int = 0,
                    extra: Optional[Dict[str, Any]] = None, include_rng: bool = True) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    payload = {
        "format": "fluxdist-ckpt-v1",
        "time": time.time(),
        "step": int(step),
        "model": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "opt_state": _cpu(opt_state),
        "extra": extra or {},
    }
    if include_rng:
        payload["rng"] = {"torch": torch.get_rng_state()}
    d = os.path.dirname(os.path.abspath(path))
    fd, tmp = tempfile.mkstemp(dir=d, prefix=".ckpt-", suffix=".tmp")
    os.close(fd)
    try:
        torch.save(payload, tmp)
        os.replace(tmp, path)  # atomic on POSIX
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return path


def load_checkpoint(path: str, model: Optional[torch.nn.Module] = None, map_location="cpu",
                    restore_rng: bool = False) -> Dict[str, Any]:
    payload = torch.load(path, map_location=map_location, weights_only=False)
    if payload.get("format") != "fluxdist-ckpt-v1":
        raise ValueError(f"{path} is not a fluxdist checkpoint")
    if model is not None:
        model.load_state_dict(payload["model"])
    if restore_rng and "rng" in payload:
        torch.set_rng_state(payload["rng"]["torch"])
    return payload


def cycle_checkpoint_name(root: str, worker: int, model_name: str, cycle: int) -> str:
    """``weights/<p>/<model>_cycle_<n>_<time>.pt`` — the reference's naming
    (``sync.jl:159``) without the hard-coded ``resnet_50`` (Q11)."""
    stamp = time.strftime("%Y-%m-%dT%H-%M-%S")
    return os.path.join(root, str(worker), f"{model_name}_cycle_{cycle}_{stamp}.pt")
