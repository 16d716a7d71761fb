"""CIFAR helpers (reference C23, ``src/cifar.jl:1-39``, dead code there).

``assemble(arrays)`` stacks per-sample arrays along a new batch dimension
(``cifar.jl:13-21``).  ``cifar_minibatch`` produces CIFAR-10-shaped synthetic
batches for the "tiny ConvNet on CIFAR-10-shaped synthetic data" config."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .synthetic import SyntheticKey, synthetic_minibatch


def assemble(arrays: Sequence[torch.Tensor], dim: int = 0) -> torch.Tensor:
    return torch.stack(list(arrays), dim=dim)


def cifar_key(n: int = 50000) -> SyntheticKey:
    return SyntheticKey(n, nclasses=10, shape=(3, 32, 32))


def cifar_minibatch(key: SyntheticKey, nsamples: int, onehot: bool = True,
                    generator: Optional[np.random.Generator] = None):
    return synthetic_minibatch(key, nsamples, onehot=onehot, generator=generator)
