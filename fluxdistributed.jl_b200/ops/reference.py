"""Plain PyTorch fp32 reference of every hand-written op (test oracle).

These play the role the library kernels play in the reference (Flux layers lowered to NNlibCUDA/cuDNN inside
``gradient(...)``, ``src/ddp_tasks.jl:80-84``): the GPU tests compare each sm_100a kernel against them
(``tests/test_native_gpu.py``), mirroring how the reference tests compare replicas against a single-device run
(``test/single_device.jl``)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def conv_fprop_ref(x, w, stride: int, pad: int):
    return F.conv2d(x.float(), w.float(), None, stride, pad)


def conv_dgrad_ref(dy, w, x_shape, stride: int, pad: int):
    return torch.nn.grad.conv2d_input(x_shape, w.float(), dy.float(), stride, pad)


def conv_wgrad_ref(dy, x, w_shape, stride: int, pad: int):
    return torch.nn.grad.conv2d_weight(x.float(), w_shape, dy.float(), stride, pad)


def bn_stats_ref(y) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel batch mean and biased variance over N,H,W of an NCHW tensor."""
    yf = y.float()
    return yf.mean(dim=(0, 2, 3)), yf.var(dim=(0, 2, 3), unbiased=False)


def bn_apply_ref(y, mean, var, gamma, beta, eps, residual=None, relu=True):
    yf = y.float()
    o = (yf - mean.view(1, -1, 1, 1)) * torch.rsqrt(var.view(1, -1, 1, 1) + eps) * gamma.float().view(1, -1, 1, 1) \
        + beta.float().view(1, -1, 1, 1)
    if residual is not None:
        o = o + residual.float()
    return F.relu(o) if relu else o


def bn_bwd_ref(dout, y, out, mean, var, gamma, eps, relu=True):
    """Returns (dy, dgamma, dbeta, dz) where dz is the gradient w.r.t. the
    pre-activation (also the gradient flowing to the residual branch)."""
    dz = dout.float()
    if relu:
        dz = dz * (out.float() > 0)
    rstd = torch.rsqrt(var + eps).view(1, -1, 1, 1)
    xhat = (y.float() - mean.view(1, -1, 1, 1)) * rstd
    M = y.numel() // y.shape[1]
    dbeta = dz.sum(dim=(0, 2, 3))
    dgamma = (dz * xhat).sum(dim=(0, 2, 3))
    dy = gamma.float().view(1, -1, 1, 1) * rstd * (dz - dbeta.view(1, -1, 1, 1) / M - xhat * dgamma.view(1, -1, 1, 1) / M)
    return dy, dgamma, dbeta, dz


def conv_bn_act_ref(x, w, gamma, beta, stride, pad, eps=1e-5, residual=None, relu=True):
    y = conv_fprop_ref(x, w, stride, pad)
    mean, var = bn_stats_ref(y)
    return bn_apply_ref(y, mean, var, gamma, beta, eps, residual, relu)


def softmax_ce_ref(logits, labels):
    """Returns (mean loss, dlogits) for integer labels."""
    lf = logits.float()
    lsm = F.log_softmax(lf, dim=-1)
    loss = -lsm.gather(1, labels.long().view(-1, 1)).mean()
    d = torch.softmax(lf, dim=-1)
    d[torch.arange(lf.shape[0], device=lf.device), labels.long()] -= 1
    return loss, d / lf.shape[0]


def bn_relu_pool_ref(y, gamma, beta, eps=1e-5):
    """Training-mode BN + ReLU + 3x3/stride-2/pad-1 max-pool (oracle of ``ops.bn_relu_pool``)."""
    import torch.nn.functional as F

    yf = y.float()
    mean = yf.mean(dim=(0, 2, 3))
    var = yf.var(dim=(0, 2, 3), unbiased=False)
    act = torch.relu((yf - mean[None, :, None, None]) * (torch.rsqrt(var + eps) * gamma)[None, :, None, None]
                     + beta[None, :, None, None])
    return F.max_pool2d(act, 3, 2, 1)


def bn_bwd_apply_ref(dz, y, mean, rstd, gamma):
    """BN backward for an already ReLU-masked output gradient ``dz`` (oracle of ``ops.bn_bwd_apply``):
    returns (dy, dgamma, dbeta)."""
    dzf, yf = dz.float(), y.float()
    m = yf.numel() // yf.shape[1]
    xhat = (yf - mean[None, :, None, None]) * rstd[None, :, None, None]
    dgamma = (dzf * xhat).sum(dim=(0, 2, 3))
    dbeta = dzf.sum(dim=(0, 2, 3))
    dy = (gamma * rstd)[None, :, None, None] * (dzf - dbeta[None, :, None, None] / m - xhat * dgamma[None, :, None, None] / m)
    return dy, dgamma, dbeta
