"""Synthetic ImageNet-/CIFAR-shaped batches (benchmark configs in BASELINE.json:
"synthetic data of the named shape").  ``SyntheticKey`` quacks like the DataFrame
``key`` the reference passes around (``len``, row slicing) so ``prepare_training``
shards it the same way, and ``synthetic_minibatch`` stands in for ``minibatch``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass
class SyntheticKey:
    """A virtual table of ``n`` samples with labels ``class_idx[i] = (i*7919) % nclasses``."""

    n: int
    nclasses: int = 1000
    shape: Tuple[int, int, int] = (3, 224, 224)
    rows: Optional[np.ndarray] = None  # row ids (after sharding)

    def __post_init__(self):
        if self.rows is None:
            self.rows = np.arange(self.n)

    def __len__(self):
        return len(self.rows)

    def take(self, idx) -> "SyntheticKey":
        return SyntheticKey(self.n, self.nclasses, self.shape, self.rows[np.asarray(idx)])

    def labels_of(self, rows) -> np.ndarray:
        return (np.asarray(rows) * 7919) % self.nclasses


def synthetic_minibatch(key: SyntheticKey, nsamples: int, *, onehot: bool = False, u8: bool = False,
                        layout: str = "CHW", generator: Optional[np.random.Generator] = None,
                        pin: bool = False, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    """Draw ``nsamples`` rows with replacement (same sampling rule as
    ``imagenet.jl:24``) and fabricate deterministic-per-row pixel data cheaply
    (one random bank, rolled by the row id) so the host cost stays negligible."""
    rng = generator or np.random.default_rng()
    rows = key.rows[rng.integers(0, len(key), size=nsamples)]
    c, h, w = key.shape
    shp = (nsamples, h, w, c) if (layout == "HWC" or u8) else (nsamples, c, h, w)
    dt = torch.uint8 if u8 else torch.float32
    if out is not None:
        x, y = out
    else:
        x = torch.empty(shp, dtype=dt, pin_memory=pin)
        y = torch.empty((nsamples,), dtype=torch.int64, pin_memory=pin)
    bank = _bank(int(np.prod(shp[1:])), u8)
    n_el = bank.numel() // 2
    for i, r in enumerate(rows):
        off = int(r * 9973) % n_el
        x[i].view(-1).copy_(bank[off:off + n_el])
    y.copy_(torch.from_numpy(key.labels_of(rows).astype(np.int64)))
    if onehot:
        return x, torch.nn.functional.one_hot(y, key.nclasses).float()
    return x, y


_BANKS = {}


def _bank(n_el: int, u8: bool) -> torch.Tensor:
    k = (n_el, u8)
    if k not in _BANKS:
        g = torch.Generator().manual_seed(1234)
        if u8:
            _BANKS[k] = torch.randint(0, 256, (2 * n_el,), dtype=torch.uint8, generator=g)
        else:
            _BANKS[k] = torch.randn((2 * n_el,), generator=g)
    return _BANKS[k]
