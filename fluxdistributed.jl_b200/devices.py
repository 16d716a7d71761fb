"""Device runtime shim (reference C18: ``src/utils.jl:1-18``).

The reference's ``@device! dev expr`` / ``synchronize()`` pair is what lets its
whole engine run on a machine without CUDA: device ids become plain dictionary
keys.  Here the same role is played by :class:`FakeDevice` (a CPU stand-in that
is hashable and orderable like a CUDA ordinal) and :func:`device_ctx`.

Three worlds share every code path above this module:

* CPU fake devices            -> ``fake_devices(n)``
* N virtual replicas on 1 GPU -> ``virtual_devices(n, cuda_index=0)``
* N real GPUs                 -> ``cuda_devices()``
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Iterable, List, Union

import torch


@dataclass(frozen=True, order=True)
class FakeDevice:
    """A logical replica slot that executes on the CPU."""

    index: int

    @property
    def torch_device(self) -> torch.device:
        return torch.device("cpu")

    @property
    def is_cuda(self) -> bool:
        return False

    def __str__(self) -> str:  # log names such as "val_fake:0"
        return f"fake:{self.index}"


@dataclass(frozen=True, order=True)
class CudaDevice:
    """A logical replica slot bound to CUDA ordinal ``cuda_index``.

    ``index`` is the replica id (dictionary key); several replicas may share one
    physical GPU (``virtual_devices``), which is the intent of the reference's
    1-GPU test branch (``test/single_device.jl:127-133``).
    """

    index: int
    cuda_index: int

    @property
    def torch_device(self) -> torch.device:
        return torch.device("cuda", self.cuda_index)

    @property
    def is_cuda(self) -> bool:
        return True

    def __str__(self) -> str:
        if self.index == self.cuda_index:
            return f"cuda:{self.cuda_index}"
        return f"cuda:{self.cuda_index}/r{self.index}"


Device = Union[FakeDevice, CudaDevice]


def functional() -> bool:
    """``CUDA.functional()`` analogue."""
    return torch.cuda.is_available()


def fake_devices(n: int) -> List[FakeDevice]:
    return [FakeDevice(i) for i in range(n)]


def cuda_devices(n: int | None = None) -> List[CudaDevice]:
    count = torch.cuda.device_count() if functional() else 0
    n = count if n is None else n
    if n > count:
        raise RuntimeError(f"requested {n} CUDA devices, {count} visible")
    return [CudaDevice(i, i) for i in range(n)]


def virtual_devices(n: int, cuda_index: int = 0) -> List[CudaDevice]:
    return [CudaDevice(i, cuda_index) for i in range(n)]


def devices(n: int | None = None) -> List[Device]:
    """``CUDA.devices()`` analogue that degrades to fake devices without a GPU."""
    if functional():
        return list(cuda_devices(n))
    return list(fake_devices(n if n is not None else 1))


def as_device(dev) -> Device:
    if isinstance(dev, (FakeDevice, CudaDevice)):
        return dev
    if isinstance(dev, torch.device):
        if dev.type == "cuda":
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            return CudaDevice(idx, idx)
        return FakeDevice(0)
    if isinstance(dev, int):
        return CudaDevice(dev, dev) if functional() else FakeDevice(dev)
    if isinstance(dev, str):
        return as_device(torch.device(dev))
    raise TypeError(f"cannot interpret {dev!r} as a device")


@contextlib.contextmanager
def device_ctx(dev):
    """``@device! dev begin ... end``: make ``dev`` current for the block."""
    dev = as_device(dev)
    if dev.is_cuda:
        with torch.cuda.device(dev.torch_device):
            yield dev
    else:
        yield dev


def synchronize(dev=None) -> None:
    """``synchronize()``: device sync on CUDA, no-op on fake devices."""
    if dev is not None:
        dev = as_device(dev)
        if dev.is_cuda:
            torch.cuda.synchronize(dev.torch_device)
        return
    if functional():
        torch.cuda.synchronize()


def to_device(obj, dev):
    """``gpu(x)`` analogue: move a tensor / module / pytree to ``dev``."""
    from .tree import fmap  # local import: tree imports nothing from here

    tdev = as_device(dev).torch_device
    if isinstance(obj, torch.nn.Module):
        return obj.to(tdev)
    return fmap(lambda t: t.to(tdev, non_blocking=True) if isinstance(t, torch.Tensor) else t, obj)


def to_cpu(obj):
    """``cpu(x)`` analogue."""
    from .tree import fmap

    if isinstance(obj, torch.nn.Module):
        return obj.to("cpu")
    return fmap(lambda t: t.detach().to("cpu") if isinstance(t, torch.Tensor) else t, obj)


def enable_peer_access(devs: Iterable[Device]) -> bool:
    """Report whether every CUDA pair in ``devs`` can address each other's memory.

    The reference warns that without P2P its buffer copies are staged through the
    host (``docs/src/training.md:30``).  PyTorch enables peer access lazily on the
    first cross-device copy; the native arena (``parallel/arena.py``) maps peers
    explicitly.
    """
    cu = sorted({d.cuda_index for d in map(as_device, devs) if d.is_cuda})
    ok = True
    for a in cu:
        for b in cu:
            if a != b and not torch.cuda.can_device_access_peer(a, b):
                ok = False
    return ok
