"""Per-rank training engine for the B200 path: static buffers + one CUDA graph per step.

The reference drives every step from the host: 2N task spawns, two global barriers and
>=3 device synchronisations per replica (``src/ddp_tasks.jl:203-228``).  Here the whole step —
input normalisation, forward, backward (gradients written straight into the flat arena), the
fused all-reduce + optimiser kernel with its device-side barriers — is captured ONCE into a
CUDA graph and replayed; the host only feeds the pinned staging ring and launches the graph
("CUDA streams and graphs instead of a tracing compiler").

``Trainer`` works for one replica per process (``torchrun``; process-based parallelism) and for
N replicas in one process (task-based parallelism, one ``Trainer`` per device thread sharing a
``FusedGroup``).
"""
from __future__ import annotations

import os

from typing import Callable, Optional, Sequence

import torch
from torch import nn

from ..devices import as_device, device_ctx
from ..losses import logitcrossentropy
from ..optim import Rule
from .fused import FusedGroup, FusedReplica


class GraphedStep:
    """Capture ``loss = fwd_bwd(transform(x), y); fused_update()`` for one replica."""

    def __init__(self, rep: FusedReplica, loss: Callable, x_shape: Sequence[int], x_dtype: torch.dtype,
                 y_shape: Sequence[int], y_dtype: torch.dtype, transform: Optional[Callable] = None,
                 warmup: int = 3, graph: bool = True, capture_error_mode: str = "global"):
        self.rep, self.loss, self.transform = rep, loss, transform
        tdev = rep.dev.torch_device
        with device_ctx(rep.dev):
            self.x = torch.zeros(tuple(x_shape), dtype=x_dtype, device=tdev)
            self.y = torch.zeros(tuple(y_shape), dtype=y_dtype, device=tdev)
            self.out = torch.zeros((), dtype=torch.float32, device=tdev)
        self.graph = None
        self.use_graph = graph and rep.dev.is_cuda
        self.warmup = warmup
        self.capture_error_mode = capture_error_mode  # "thread_local" when several replica threads capture at once
        self.restore_after_warmup = True
        self.launches_per_step = 0
        self._captured = False

    def _body(self):
        x = self.transform(self.x) if self.transform is not None else self.x
        l = self.rep.forward_backward(x, self.y, self.loss)
        self.rep.reduce_and_update()
        self.out.copy_(l.float())

    def capture(self):
        """Eager warm-up steps, then capture.  The warm-up runs real steps on whatever sits in the static buffers (lazy
        module loading, cuDNN/TMA descriptor caches, allocator pools all have to be warm before a capture), so the
        replica's weights, optimiser moments, step counters and BN running statistics are snapshotted before and
        restored after: capturing has no training side effect and the first replay is exactly step 1 (the
        first batch used to be applied three times).  In multi-replica groups every replica does the same, in step."""
        from .. import ops

        rep = self.rep
        with device_ctx(rep.dev):
            if not self.use_graph:
                self._captured = True
                return self
            snap = rep.snapshot_state() if self.restore_after_warmup else None
            # high priority: the step's critical path (forward, BN backward, dgrad) is issued on this stream; the side
            # stream of the weight-gradient kernels keeps the default (lowest) priority, so when both have CTAs
            # pending the block scheduler places the critical-path ones first
            s = torch.cuda.Stream(device=rep.dev.torch_device, priority=int(os.environ.get("FLUXDIST_MAIN_PRIO", "-1")))
            s.wait_stream(torch.cuda.current_stream(rep.dev.torch_device))
            with torch.cuda.stream(s):
                for _ in range(self.warmup):
                    self._body()
            torch.cuda.current_stream(rep.dev.torch_device).wait_stream(s)
            torch.cuda.synchronize(rep.dev.torch_device)
            n0 = ops.extension().launch_count() if ops.native_available() else 0
            self.graph = torch.cuda.CUDAGraph()
            # an explicit per-replica capture stream: torch's default capture stream is a process-wide singleton bound
            # to whichever device used it first, which breaks concurrent captures from several replica threads
            with torch.cuda.graph(self.graph, stream=s, capture_error_mode=self.capture_error_mode):
                self._body()
            self.launches_per_step = (ops.extension().launch_count() - n0) if ops.native_available() else 0
            if snap is not None:
                torch.cuda.synchronize(rep.dev.torch_device)
                rep.restore_state(snap)
            self._captured = True
        return self

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run one step.  ``x``/``y`` (device tensors) are copied into the static buffers on the
        current stream; pass None when the caller already filled ``self.x``/``self.y``."""
        if not self._captured:
            self.capture()
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            from .. import ops

            n0 = ops.extension().launch_count() if ops.native_available() else 0
            self._body()
            if ops.native_available():
                self.launches_per_step = ops.extension().launch_count() - n0
        return self.out


class Trainer:
    """One replica's trainer: model + fused engine + graphed step + optional staged input.

    ``input="u8"``: batches arrive as uint8 NHWC (what the pinned ring carries; 4x fewer PCIe
    bytes than the reference's fp32 upload) and are normalised on the GPU; ``"float"``: float
    NCHW batches are used as they are.
    """

    def __init__(self, model: nn.Module, opt: Rule, device, batch: int, *, loss: Callable = logitcrossentropy,
                 image_hw: int = 224, input: str = "u8", graph: bool = True, comm: str = "auto",
                 backend: str = "auto", dtype: Optional[torch.dtype] = None, group=None,
                 replica: Optional[FusedReplica] = None, mode: Optional[int] = None):
        from ..nn import set_backend

        self.dev = as_device(device)
        set_backend(backend)
        tdev = self.dev.torch_device
        with device_ctx(self.dev):
            if replica is None:
                model = model.to(tdev)
                if any(p.dim() == 4 for p in model.parameters()):
                    model = model.to(memory_format=torch.channels_last)
                kw = {} if mode is None else {"mode": mode}
                self.group = FusedGroup.for_process(model, self.dev, opt, loss=loss, dtype=dtype, comm=comm,
                                                    group=group, **kw)
                self.rep = self.group.replicas[0]
            else:
                self.rep, self.group = replica, replica.group
        self.model = self.rep.model
        self.batch, self.hw, self.input = batch, image_hw, input
        lowp = self.rep.arena.w16.dtype if self.rep.arena.w16 is not None else torch.float32
        if input == "u8":
            x_shape, x_dtype = (batch, image_hw, image_hw, 3), torch.uint8
            transform = self._normalize_u8(lowp)
        else:
            x_shape, x_dtype = (batch, 3, image_hw, image_hw), torch.float32
            transform = (lambda x: x.to(lowp).contiguous(memory_format=torch.channels_last)) if self.dev.is_cuda else None
        self.step_fn = GraphedStep(self.rep, loss, x_shape, x_dtype, (batch,), torch.int64, transform, graph=graph)

    def _normalize_u8(self, lowp):
        from .. import ops
        from ..data import MEAN, STD

        if self.dev.is_cuda and ops.native_available() and lowp == torch.bfloat16:
            return lambda x: ops.normalize_u8(x, 3)
        mean = torch.tensor(MEAN, device=self.dev.torch_device).view(1, 1, 1, 3)
        std = torch.tensor(STD, device=self.dev.torch_device).view(1, 1, 1, 3)
        return lambda x: ((x.float() / 255.0 - mean) / std).permute(0, 3, 1, 2).to(lowp)

    def capture(self):
        self.step_fn.capture()
        return self

    def step(self, x=None, y=None) -> torch.Tensor:
        return self.step_fn(x, y)

    @property
    def launches_per_step(self) -> int:
        return self.step_fn.launches_per_step

    def status(self):
        c = self.group.comm
        return c.status(self.rep) if hasattr(c, "status") else (False, False)

    def finish(self):
        self.group.sync_to_modules()
        return self.model


def graphed_replica_step(rep: FusedReplica, x: torch.Tensor, y: torch.Tensor, loss: Callable) -> torch.Tensor:
    """Task mode (one thread per GPU in one process): run ``rep``'s step through a per-replica CUDA graph keyed by
    the batch signature.  The first call performs the eager warm-up steps (real training steps, all replicas take
    part in their barriers) and captures with thread-local capture mode; later calls only copy the batch into the
    static buffers and replay."""
    key = (tuple(x.shape), x.dtype, tuple(y.shape), y.dtype)
    cache = rep.__dict__.setdefault("_graphs", {})
    gs = cache.get(key)
    if gs is None:
        lowp = rep.arena.w16.dtype if rep.arena.w16 is not None else torch.float32
        if x.dtype == torch.uint8:
            from .. import ops

            transform = (lambda t: ops.normalize_u8(t, 3))
        elif x.dim() == 4 and rep.dev.is_cuda:
            transform = (lambda t: t.to(lowp).contiguous(memory_format=torch.channels_last))
        else:
            transform = None
        gs = GraphedStep(rep, loss, x.shape, x.dtype, y.shape, y.dtype, transform, warmup=2,
                         capture_error_mode="thread_local")
        with device_ctx(rep.dev):
            gs.x.copy_(x)
            gs.y.copy_(y)
        gs.capture()
        cache[key] = gs
    with device_ctx(rep.dev):
        return gs(x, y)
