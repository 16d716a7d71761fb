"""Legacy "train independently, keep the arg-min-loss model" strategy
(reference C22, dead code in ``src/test.jl:1-63``), kept for capability parity.

Per cycle every worker trains its own copy on its own sample draw
(``distribute``, ``:26-41``); the copy with the lowest validation loss becomes
everybody's starting point for the next cycle (``findmin``, ``:58``); the learning
rate is divided by 5 every 10 cycles (``:50``); the cycle time is logged (``:52,60``).
"""
from __future__ import annotations

import copy
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import nn

from ..devices import as_device, device_ctx
from ..optim import ADAM, Rule, state as opt_state
from ..tree import gradient
from ..utils.loggers import get_logger


def main(loss, model: nn.Module, data, opt: Rule, batchsize: int = 48, epochs: int = 1, device=None):
    """Plain single-device epoch loop (``src/test.jl:1-11``)."""
    from ..data import DataLoader

    dev = as_device(device) if device is not None else as_device("cpu")
    st = opt_state(opt, model)
    with device_ctx(dev):
        model = model.to(dev.torch_device)
        for b in DataLoader(lambda b: tuple(t.to(dev.torch_device) for t in b), data, batchsize=batchsize,
                            epochs=epochs):
            x, y = b
            gs, _ = gradient(lambda m: loss(m(x), y), model)
            model, st = opt(model, gs, st)
    return model


def tmp(loss, model: nn.Module, data, val, opt: Rule, batchsize: int, device) -> Tuple[nn.Module, float]:
    """One worker's independent training; returns ``(cpu_model, val_loss)`` (``:13-24``)."""
    dev = as_device(device)
    m = main(loss, copy.deepcopy(model), data, opt, batchsize, 1, dev)
    with device_ctx(dev), torch.no_grad():
        vl = float(loss(m(val[0].to(dev.torch_device)), val[1].to(dev.torch_device)))
    return m.to("cpu"), vl


def distribute(loss, model, draw: Callable[[int], tuple], val, opt, devices: Sequence, batchsize: int = 48):
    """Every worker trains on its own draw (``:26-41``).  Unlike the reference (Q16)
    workers use *their own* device, not all ``device!(0)``."""
    with ThreadPoolExecutor(max_workers=len(devices)) as ex:
        futs = [ex.submit(tmp, loss, model, draw(i), val, opt, batchsize, d) for i, d in enumerate(devices)]
        return [f.result() for f in futs]


def run_distributed_argmin(loss, model, draw, val, devices, cycles: int = 1, opt: Optional[Rule] = None,
                           batchsize: int = 48, lr_drop_every: int = 10, lr_drop: float = 5.0, logger=None):
    """``run_distributed`` of ``src/test.jl:43-63``."""
    lg = logger or get_logger()
    opt = opt or ADAM(1e-3, (0.9, 0.99))
    for n in range(1, cycles + 1):
        if n % lr_drop_every == 0:
            opt = opt.with_eta(opt.eta / lr_drop)
        t0 = time.perf_counter()
        res = distribute(loss, model, draw, val, opt, devices, batchsize)
        losses = [r[1] for r in res]
        best = min(range(len(res)), key=lambda i: losses[i])
        model = res[best][0]
        lg.info("cycle", n=n, time=time.perf_counter() - t0, best=best, val_loss=losses[best])
    return model
