"""Layers.

The framework layer is PyTorch: plain ``torch.nn`` modules (``Conv2d``, ``Linear``,
``BatchNorm2d``, ``MaxPool2d`` …) work everywhere in this package, mirroring the
Flux layers the reference trains (E2 in SURVEY §1).  This module adds the *fused*
building block of the model zoo:

``ConvBN`` = convolution (no bias) -> BatchNorm -> (+ residual) -> (ReLU)

which is the unit the sm_100a path executes as
``conv_fprop (tcgen05 implicit GEMM)`` -> ``bn_stats`` -> ``bn_apply(+res,+relu)`` and
whose backward is ``bn_bwd_reduce`` -> ``bn_bwd_apply`` -> ``conv_dgrad`` /
``conv_wgrad``.  On CPU, or when the extension is not built,
the same module runs the PyTorch reference ops, which is also the numerics oracle
for the kernels.

Tensors are logically NCHW with ``torch.channels_last`` strides (physically NHWC);
conv weights are ``[K,C,R,S]`` channels_last (physically KRSC) — the layouts the
kernels consume without any repacking.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

_BACKEND = {"mode": "auto"}  # "auto" | "torch" | "native"


def set_backend(mode: str) -> None:
    """Select the op backend for ``ConvBN``/``Linear`` heads: ``auto`` uses the sm_100a
    kernels on CUDA tensors when the extension is loaded, else PyTorch ops."""
    assert mode in ("auto", "torch", "native")
    _BACKEND["mode"] = mode


def get_backend() -> str:
    return _BACKEND["mode"]


def _use_native(x: torch.Tensor) -> bool:
    mode = _BACKEND["mode"]
    if mode == "torch" or not x.is_cuda:
        return False
    from .. import ops

    if mode == "native":
        ops.require_native()
        return True
    return ops.native_available() and x.dtype == torch.bfloat16


class _BNParams(nn.Module):
    """Holds BatchNorm affine parameters and running statistics."""

    def __init__(self, c: int, eps: float = 1e-5, momentum: float = 0.1):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.eps, self.momentum = eps, momentum

    def forward(self, x):  # usable stand-alone as an ordinary BatchNorm2d
        return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                            self.training, self.momentum, self.eps)


class _ConvParams(nn.Module):
    def __init__(self, cin: int, cout: int, k: int):
        super().__init__()
        w = torch.empty(cout, cin, k, k)
        nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")
        self.weight = nn.Parameter(w.contiguous(memory_format=torch.channels_last))


class ConvBN(nn.Module):
    def __init__(self, cin: int, cout: int, k: int, stride: int = 1, padding: Optional[int] = None,
                 relu: bool = True, eps: float = 1e-5, momentum: float = 0.1, zero_init: bool = False):
        super().__init__()
        self.cin, self.cout, self.k, self.stride = cin, cout, k, stride
        self.padding = (k // 2) if padding is None else padding
        self.relu = relu
        self.conv = _ConvParams(cin, cout, k)
        self.bn = _BNParams(cout, eps, momentum)
        if zero_init:
            nn.init.zeros_(self.bn.weight)

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None, add_link=None, res_link=None,
                park_link=None, pool: bool = False, own_link=None, src_link=None) -> torch.Tensor:
        if _use_native(x):
            from ..ops import conv_bn_act
            from ..ops.native import bn_workspace

            ws = self.__dict__.get("_bn_ws")
            if ws is None or ws.device != x.device:
                ws = self.__dict__["_bn_ws"] = bn_workspace(self.cout, x.device)
            return conv_bn_act(x, self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean,
                               self.bn.running_var, residual, self.relu, self.stride, self.padding,
                               self.bn.eps, self.bn.momentum, self.training,
                               getattr(self.conv.weight, "_lowp", None),
                               (getattr(self.conv.weight, "_gradbuf", None), getattr(self.bn.weight, "_gradbuf", None),
                                getattr(self.bn.bias, "_gradbuf", None)), ws,
                               add_link, res_link, park_link, pool, own_link, src_link)
        w = self.conv.weight
        if x.dtype != w.dtype:
            w = w.to(x.dtype)
        y = F.conv2d(x, w, None, self.stride, self.padding)
        y = F.batch_norm(y, self.bn.running_mean, self.bn.running_var, self.bn.weight, self.bn.bias,
                         self.training, self.bn.momentum, self.bn.eps)
        if residual is not None:
            y = y + residual
        y = F.relu(y) if self.relu else y
        return F.max_pool2d(y, 3, 2, 1) if pool else y

    def extra_repr(self):
        return f"{self.cin}, {self.cout}, k={self.k}, s={self.stride}, p={self.padding}, relu={self.relu}"


class Head(nn.Module):
    """Global average pool + flatten + Linear (K7/K8: GAP fused into the FC prologue)."""

    def __init__(self, cin: int, nclasses: int):
        super().__init__()
        self.fc = nn.Linear(cin, nclasses)

    def forward(self, x):
        if _use_native(x):
            from ..ops import gap_linear

            return gap_linear(x, self.fc.weight, self.fc.bias,
                              getattr(self.fc.weight, "_lowp", None))
        p = x.mean(dim=(2, 3))
        w, b = self.fc.weight, self.fc.bias
        if p.dtype != w.dtype:
            w, b = w.to(p.dtype), b.to(p.dtype)
        return F.linear(p, w, b)


class MaxPool(nn.Module):
    def __init__(self, k: int = 3, stride: int = 2, padding: int = 1):
        super().__init__()
        self.k, self.stride, self.padding = k, stride, padding

    def forward(self, x):
        if _use_native(x):
            from ..ops import maxpool2d

            return maxpool2d(x, self.k, self.stride, self.padding)
        return F.max_pool2d(x, self.k, self.stride, self.padding)


class Flatten(nn.Flatten):
    """``Flux.flatten``."""


class Conv2dN(nn.Conv2d):
    """``nn.Conv2d`` whose CUDA/bf16 path runs this repository's kernels: the tcgen05 implicit GEMM when both channel
    counts are multiples of 64 (no bias), otherwise the direct CUDA-core convolution (``csrc/direct_conv.cu``) — so
    the reference's ``Conv((7,7), 3=>3)`` test model no longer falls back to cuDNN.  Same parameters / state dict as
    ``nn.Conv2d``; CPU and fp32 inputs take the PyTorch path."""

    def forward(self, x):
        simple = (self.groups == 1 and self.dilation == (1, 1) and self.stride[0] == self.stride[1]
                  and self.padding[0] == self.padding[1] and isinstance(self.padding, tuple) and x.dim() == 4)
        if simple and _use_native(x) and x.dtype == torch.bfloat16:
            from ..ops.native import direct_conv2d

            return direct_conv2d(x, self.weight, self.bias, self.stride[0], self.padding[0])
        w, b = self.weight, self.bias
        if x.dtype != w.dtype:
            w = w.to(x.dtype)
            b = b.to(x.dtype) if b is not None else None
        return F.conv2d(x, w, b, self.stride, self.padding, self.dilation, self.groups)


class LinearN(nn.Linear):
    """``nn.Linear`` on this repository's GEMM kernels for CUDA/bf16 inputs (tcgen05 or the small-GEMM kernel)."""

    def forward(self, x):
        if x.dim() == 2 and _use_native(x) and x.dtype == torch.bfloat16:
            from ..ops.native import linear

            return linear(x, self.weight, self.bias, getattr(self.weight, "_lowp", None))
        w, b = self.weight, self.bias
        if x.dtype != w.dtype:
            w = w.to(x.dtype)
            b = b.to(x.dtype) if b is not None else None
        return F.linear(x, w, b)


def Chain(*layers: nn.Module) -> nn.Sequential:
    """``Flux.Chain``."""
    return nn.Sequential(*layers)


def Dense(cin: int, cout: int, bias: bool = True) -> nn.Linear:
    """``Flux.Dense``."""
    return LinearN(cin, cout, bias=bias)


def Conv(k, ch, stride: int = 1, pad: int = 0, bias: bool = True) -> nn.Conv2d:
    """``Flux.Conv((k,k), cin=>cout)``; ``ch`` is ``(cin, cout)``."""
    k = k if isinstance(k, int) else k[0]
    return Conv2dN(ch[0], ch[1], k, stride=stride, padding=pad, bias=bias)


def BatchNorm(c: int) -> nn.BatchNorm2d:
    return nn.BatchNorm2d(c)
