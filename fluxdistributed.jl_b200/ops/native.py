"""Python entry points of the sm_100a kernels (``csrc/``) and the autograd glue.

Conventions: activations are logically NCHW with ``torch.channels_last`` strides
(physically NHWC), conv weights logically ``[K,C,R,S]`` channels_last (physically
KRSC), everything bf16 with fp32 statistics/gradients.  Each function has a plain
PyTorch oracle in ``ops/reference.py``; ``tests/test_native_gpu.py`` compares them.

Reference call sites these replace: the model forward/backward inside ``gradient(m -> loss(m(x), y), m)``
(``src/ddp_tasks.jl:80-84``, ``src/sync.jl:140-142``), which the reference leaves to Flux/NNlibCUDA/cuDNN —
K1-K3 convolutions (``conv_fprop`` / ``conv_dgrad`` / ``conv_wgrad``, halo and stem variants), K4-K6 BatchNorm and
activation (``bn_apply``, ``bn_bwd_fused``, ``bn_bwd_apply``, statistics in the conv epilogues), K7 pooling
(``maxpool2d``, ``bn_relu_pool``, GAP inside ``gap_linear``), K8 the Dense head (``gap_linear``), K9 the loss
(``softmax_ce``; reference ``loss`` at ``src/ddp_tasks.jl:28``), K15/K20 the input path (``normalize_u8``; reference
``Flux.normalise`` per image at ``src/imagenet.jl:31-34``).  The autograd glue (``_ConvBNAct`` and friends) plays the
role of Zygote's pullbacks for those layers.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _ext():
    from . import extension

    return extension()


# --------------------------------------------------------------------------- layout helpers
def nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical NCHW (channels_last) -> an NHWC-shaped contiguous view."""
    v = x.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


def from_nhwc(v: torch.Tensor) -> torch.Tensor:
    """NHWC-shaped contiguous tensor -> logical NCHW channels_last view."""
    return v.permute(0, 3, 1, 2)


def krsc(w: torch.Tensor) -> torch.Tensor:
    v = w.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


# --------------------------------------------------------------------------- raw kernels
def gemm_nn(a: torch.Tensor, b: torch.Tensor, out_dtype=BF16) -> torch.Tensor:
    """``a[M,K] @ b[K,N]`` on tcgen05 (B read N-major)."""
    d = torch.empty(a.shape[0], b.shape[1], dtype=out_dtype, device=a.device)
    _ext().gemm_nn(a.contiguous(), b.contiguous(), d)
    return d


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
    """``a[K,M]^T @ b[K,N]`` on tcgen05 (both operands MN-major)."""
    d = torch.empty(a.shape[1], b.shape[1], dtype=out_dtype, device=a.device)
    _ext().gemm_tn(a.contiguous(), b.contiguous(), d)
    return d


def gemm_nt(a: torch.Tensor, b: torch.Tensor, out_dtype=BF16) -> torch.Tensor:
    """``a[M,K] @ b[N,K]^T`` on tcgen05 (bf16 in, fp32 accumulate)."""
    d = torch.empty(a.shape[0], b.shape[0], dtype=out_dtype, device=a.device)
    _ext().gemm_nt(a.contiguous(), b.contiguous(), d)
    return d


def conv_fprop(x: torch.Tensor, w: torch.Tensor, stride: int, pad: int, stats=None):
    """``x`` NCHW/channels_last bf16, ``w`` [K,C,R,S]/channels_last bf16 -> y (same convention).
    ``stats``: ``True`` -> also return per-channel (sum, sum of squares) of the fp32 accumulators, reduced in the
    epilogue (fused BN statistics); a ``(sum, sumsq)`` pair of ZEROED fp32 buffers -> accumulate into them."""
    xh, wk = nhwc(x), krsc(w)
    n, h, ww, _ = xh.shape
    k, r, s, _ = wk.shape
    p, q = (h + 2 * pad - r) // stride + 1, (ww + 2 * pad - s) // stride + 1
    y = torch.empty(n, p, q, k, dtype=BF16, device=x.device)
    if stats is None or stats is False:
        _ext().conv_fprop(xh, wk, y, stride, pad, None, None)
        return from_nhwc(y)
    sm, sq = (torch.zeros(k, dtype=torch.float32, device=x.device), torch.zeros(k, dtype=torch.float32, device=x.device)) \
        if stats is True else stats
    _ext().conv_fprop(xh, wk, y, stride, pad, sm, sq)
    return from_nhwc(y), sm, sq


def conv3x3_fprop(x: torch.Tensor, w: torch.Tensor, stats=False, variant: int = 0):
    """Halo-tile 3x3/stride-1/pad-1 forward conv; with ``stats`` also returns the per-channel
    (sum, sum of squares) of the bf16 output accumulated in the epilogue (fused BN statistics).
    ``stats`` may be a ``(sum, sumsq)`` pair of ZEROED buffers to accumulate into (no memset is issued)."""
    xh, wk = nhwc(x), krsc(w)
    n, h, ww, _ = xh.shape
    k = wk.shape[0]
    y = torch.empty(n, h, ww, k, dtype=BF16, device=x.device)
    s = ss = None
    if isinstance(stats, tuple):
        s, ss = stats
        variant |= 0x100
    elif stats:
        s = torch.empty(k, dtype=torch.float32, device=x.device)
        ss = torch.empty(k, dtype=torch.float32, device=x.device)
    _ext().conv3x3_fprop(xh, wk, y, s, ss, variant)
    return (from_nhwc(y), s, ss) if stats else from_nhwc(y)


def _bn_epi_list(bn_epi, device):
    y_, out_, mean_, rstd_, gamma_, beta_, s0_, s1_ = bn_epi
    return [nhwc(y_), nhwc(out_) if out_ is not None else torch.empty(0, dtype=BF16, device=device), mean_, rstd_,
            gamma_, beta_, s0_, s1_]


def conv3x3_dgrad(dy: torch.Tensor, w: torch.Tensor, variant: int = 0, add: Optional[torch.Tensor] = None,
                  bn_epi=None) -> torch.Tensor:
    """``add`` (same shape as the result) is summed into dx in the epilogue (fused residual-gradient accumulation).
    ``bn_epi``: dz epilogue as in :func:`conv_dgrad` (ReLU mask + BN-backward sums of the producing BN layer)."""
    dyh, wk = nhwc(dy), krsc(w)
    n, h, ww, _ = dyh.shape
    dx = torch.empty(n, h, ww, wk.shape[3], dtype=BF16, device=dy.device)
    epi = None if bn_epi is None else _bn_epi_list(bn_epi, dy.device)
    if epi is not None and variant == 0:
        variant = 1 if wk.shape[3] == 64 and wk.shape[0] == 64 else 3
    _ext().conv3x3_dgrad(dyh, wk, dx, variant, None if add is None else nhwc(add), epi)
    return from_nhwc(dx)


def conv3x3_wgrad(dy: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None, splits: int = 0):
    dyh, xh = nhwc(dy), nhwc(x)
    k, c = dyh.shape[3], xh.shape[3]
    if out is None:
        buf = torch.zeros(k, 3, 3, c, dtype=torch.float32, device=dy.device)
        _ext().conv3x3_wgrad(dyh, xh, buf, splits)
        return buf.permute(0, 3, 1, 2)
    buf = out.permute(0, 2, 3, 1)
    assert buf.is_contiguous() and buf.dtype == torch.float32
    _ext().conv3x3_wgrad(dyh, xh, buf, splits)
    return out


def stem_s2d(x: torch.Tensor) -> torch.Tensor:
    """bf16 NHWC-physical image batch [N,3,H,W] -> space-to-depth tensor X2 [N,(H+6)/2,(W+6)/2,16]."""
    xh = nhwc(x)
    n, h, w, c = xh.shape
    assert c == 3 and h % 2 == 0 and w % 2 == 0
    x2 = torch.empty(n, (h + 6) // 2, (w + 6) // 2, 16, dtype=BF16, device=x.device)
    _ext().s2d_pad(xh, x2, 3)
    return x2


def stem_fprop(x2: torch.Tensor, w: torch.Tensor, stats: bool = False):
    """7x7/2/pad-3 conv, 3->64 channels, as a 4x4/1 tcgen05 conv over the s2d input ``x2``."""
    wk = krsc(w)
    w2 = torch.empty(wk.shape[0], 256, dtype=BF16, device=w.device)
    _ext().stem_w_s2d(wk, w2)
    n, hs, ws, _ = x2.shape
    y = torch.empty(n, hs - 3, ws - 3, 64, dtype=BF16, device=x2.device)
    s = ss = None
    pre = isinstance(stats, tuple)
    if pre:
        s, ss = stats
    elif stats:
        s = torch.empty(64, dtype=torch.float32, device=x2.device)
        ss = torch.empty(64, dtype=torch.float32, device=x2.device)
    _ext().stem_fprop(x2, w2, y, s, ss, pre)
    return (from_nhwc(y), s, ss) if stats else from_nhwc(y)


def stem_wgrad(dy: torch.Tensor, x2: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Weight gradient of the stem conv as [64,3,7,7] (channels_last) fp32, accumulated into ``out`` if given."""
    dw2 = torch.zeros(64, 256, dtype=torch.float32, device=dy.device)
    _ext().stem_wgrad(nhwc(dy), x2, dw2)
    if out is None:
        out = torch.zeros(64, 3, 7, 7, dtype=torch.float32, device=dy.device).contiguous(memory_format=torch.channels_last)
    buf = out.permute(0, 2, 3, 1)
    assert buf.is_contiguous()
    _ext().stem_dw_from_s2d(dw2, buf)
    return out


S2_NATIVE = os.environ.get("FLUXDIST_S2_DGRAD", "1") != "0"


def _s2_native_ok(h: int, ww: int, r: int, s: int, pad: int) -> bool:
    return (S2_NATIVE and r <= 3 and s <= 3 and h % 2 == 0 and ww % 2 == 0
            and (h + 2 * pad - r) // 2 + 1 == h // 2 and (ww + 2 * pad - s) // 2 + 1 == ww // 2)


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, x_shape, stride: int, pad: int,
               add: Optional[torch.Tensor] = None, add_compact: Optional[torch.Tensor] = None,
               bn_epi=None) -> torch.Tensor:
    """Data gradient.  Stride 1 runs the tcgen05 dgrad directly.  Stride 2 (1x1 / 3x3 filters, even images) runs the
    parity-decomposed kernel: each (h%2, w%2) class of dx is a stride-1 correlation of dy with the taps that reach it,
    so neither a zero-upsampled copy of dy nor multiplications by its zeros exist.  ``add_compact`` ([N,C,H/2,W/2]) is
    summed into the (even, even) class: the data gradient of a parallel 1x1/stride-2 projection (ResNet downsample).
    Other strided geometries zero-upsample ``dy`` and reuse the stride-1 kernel.
    ``bn_epi`` (stride 1 only) = ``(y, out|None, mean, rstd, gamma, beta, s0, s1)`` of the BN(+ReLU) layer whose output
    gradient this is: the epilogue then stores the ReLU-masked gradient dz and adds sum dz*(y-mean) / sum dz into
    ``s0`` / ``s1`` (the reduction pass of that layer's backward; finish with :func:`bn_bwd_apply`)."""
    n, c, h, ww = x_shape
    dyh, wk = nhwc(dy), krsc(w)
    k, r, s, _ = wk.shape
    if stride == 2 and add is None and _s2_native_ok(h, ww, r, s, pad):
        full = all(any((p + pad - t) % 2 == 0 for t in range(r)) for p in (0, 1)) and \
            all(any((p + pad - t) % 2 == 0 for t in range(s)) for p in (0, 1))
        alloc = torch.empty if full else torch.zeros  # classes no tap reaches (1x1 filters) stay zero
        dx = alloc(n, h, ww, c, dtype=BF16, device=dy.device)
        _ext().conv_dgrad_s2(dyh, wk, dx, pad, None if add_compact is None else nhwc(add_compact))
        return from_nhwc(dx)
    if add_compact is not None:
        raise ValueError("add_compact needs the native stride-2 dgrad")
    if stride != 1:
        hu, wu = h + 2 * pad - r + 1, ww + 2 * pad - s + 1
        up = torch.empty(n, hu, wu, k, dtype=BF16, device=dy.device)
        _ext().upsample_zero(dyh, up, stride)
        dyh = up
    dx = torch.empty(n, h, ww, c, dtype=BF16, device=dy.device)
    epi = None
    if bn_epi is not None:
        assert stride == 1, "the dz epilogue is implemented for stride-1 dgrad"
        y_, out_, mean_, rstd_, gamma_, beta_, s0_, s1_ = bn_epi
        epi = [nhwc(y_), nhwc(out_) if out_ is not None else torch.empty(0, dtype=BF16, device=dy.device), mean_, rstd_,
               gamma_, beta_, s0_, s1_]
    _ext().conv_dgrad(dyh, wk, dx, pad, None if add is None else nhwc(add), epi)
    return from_nhwc(dx)


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, w_shape, stride: int, pad: int,
               out: Optional[torch.Tensor] = None, splits: int = 0, accumulate: bool = True) -> torch.Tensor:
    """Weight gradient in fp32, accumulated into ``out`` (a [K,C,R,S] channels_last view of the
    gradient arena) when given.  ``accumulate=False`` overwrites instead (only honoured without split-K)."""
    k, c, r, s = w_shape
    dyh, xh = nhwc(dy), nhwc(x)
    if out is None:
        buf = torch.zeros(k, r, s, c, dtype=torch.float32, device=dy.device)
        _ext().conv_wgrad(dyh, xh, buf, stride, pad, splits, True)
        return buf.permute(0, 3, 1, 2)
    buf = out.permute(0, 2, 3, 1)
    assert buf.is_contiguous() and buf.dtype == torch.float32
    _ext().conv_wgrad(dyh, xh, buf, stride, pad, splits, accumulate)
    return out


def bn_stats(y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-channel (sum, sum of squares) of an NCHW/channels_last bf16 tensor."""
    c = y.shape[1]
    s = torch.empty(c, dtype=torch.float32, device=y.device)
    ss = torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_stats(nhwc(y), s, ss)
    return s, ss


def bn_apply(y, s, ss, gamma, beta, eps, residual=None, relu=True, running_mean=None, running_var=None,
             momentum=0.1, training=True, consumed=None):
    """``consumed``: int32 device counter; when given, the last CTA to have read ``s``/``ss`` zeroes them (they are
    epilogue-accumulated statistics buffers that the producing conv adds into on its next launch)."""
    yh = nhwc(y)
    out = torch.empty_like(yh)
    c = y.shape[1]
    mean = torch.empty(c, dtype=torch.float32, device=y.device)
    rstd = torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_apply(yh, None if residual is None else nhwc(residual), out, s, ss, gamma, beta, running_mean,
                    running_var, mean, rstd, eps, momentum, training, relu, consumed)
    return from_nhwc(out), mean, rstd


def bn_bwd_apply(dz, y, mean, rstd, gamma, s0, s1, dgamma=None, dbeta=None, consumed=None):
    """Single-pass BN backward for a pre-masked gradient ``dz`` whose per-channel sums were reduced by the producing
    dgrad's epilogue (``conv_dgrad(..., bn_epi=...)``).  Returns (dy, dgamma, dbeta)."""
    yh = nhwc(y)
    c = y.shape[1]
    dy = torch.empty_like(yh)
    dgamma = dgamma if dgamma is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    dbeta = dbeta if dbeta is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_bwd_apply(nhwc(dz), yh, mean, rstd, gamma, s0, s1, dgamma, dbeta, dy, consumed)
    return from_nhwc(dy), dgamma, dbeta


def bn_relu_pool(y, s, ss, gamma, beta, eps, running_mean=None, running_var=None, momentum=0.1, training=True,
                 consumed=None):
    """BN + ReLU + 3x3/stride-2/pad-1 max-pool in one pass over the raw conv output ``y`` (the stem tail):
    returns (pooled, argmax bytes, mean, rstd).  The full-resolution BN output is never materialised."""
    yh = nhwc(y)
    n, h, w, c = yh.shape
    p, q = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty(n, p, q, c, dtype=BF16, device=y.device)
    idx = torch.empty(n, p, q, c, dtype=torch.uint8, device=y.device)
    mean = torch.empty(c, dtype=torch.float32, device=y.device)
    rstd = torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_relu_pool(yh, out, idx, s, ss, gamma, beta, running_mean, running_var, mean, rstd, eps, momentum, training,
                        consumed)
    return from_nhwc(out), idx, mean, rstd


def bn_bwd(dout, y, out, mean, rstd, gamma, relu=True, want_dres=False, dgamma=None, dbeta=None, beta=None):
    yh = nhwc(y)
    c = y.shape[1]
    dy = torch.empty_like(yh)
    dres = torch.empty_like(yh) if want_dres else None
    dgamma = dgamma if dgamma is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    dbeta = dbeta if dbeta is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    # `out=None` with relu: the mask is recomputed from y (valid when there is no residual input)
    beta = beta if beta is not None else torch.zeros_like(gamma)
    _ext().bn_bwd(nhwc(dout), yh, nhwc(out) if (relu and out is not None) else None, mean, rstd, gamma, beta, dgamma,
                  dbeta, dy, dres, relu)
    return from_nhwc(dy), dgamma, dbeta, (from_nhwc(dres) if want_dres else None)


import os as _os

DZ_EPILOGUE = _os.environ.get("FLUXDIST_DZ_EPILOGUE", "1") != "0"  # BN-backward reduction in the dgrad epilogue
# rows (N*H*W) up to which the dz epilogue replaces the two-pass BN backward.  With eight epilogue warps the epilogue hides
# behind the MMAs on every ResNet stage (round-2 measurement, gpurun_out/r2p1: 28x28x128 dgrad +9 us, BN backward -15 us)
DZ_MAX_ROWS = int(_os.environ.get("FLUXDIST_DZ_MAX_ROWS", "400000"))
# chain the dz epilogue across blocks: the next block's first conv masks and reduces for this block's LAST BN layer
# (mask from the post-residual activation).  Validated on B200 in round 2 (tests + 18.5k -> 19.2k img/s with EPI8).
DZ_CHAIN = _os.environ.get("FLUXDIST_DZ_CHAIN", "1") == "1"
EPI_STATS = _os.environ.get("FLUXDIST_EPI_STATS", "1") != "0"  # BN statistics in the conv epilogues
SIDE_WGRAD = _os.environ.get("FLUXDIST_SIDE_WGRAD", "1") != "0"  # wgrad kernels on a side stream (off the critical path)
_SIDE = {}
# CUDA stream priority of the weight-gradient side stream (0 = lowest; the step's capture stream uses FLUXDIST_MAIN_PRIO,
# default -1: when both streams have CTAs pending, the critical path goes first)
SIDE_PRIORITY = int(_os.environ.get("FLUXDIST_SIDE_PRIO", "0"))
_SIDE_PENDING = set()


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    key = (device.index if device.index is not None else torch.cuda.current_device())
    s = _SIDE.get(key)
    if s is None:
        s = _SIDE[key] = torch.cuda.Stream(device=device, priority=SIDE_PRIORITY)
    return s


def join_side_stream(device=None) -> None:
    """Make the current stream wait for weight-gradient work queued on the side stream."""
    if not _SIDE_PENDING:
        return
    for key in list(_SIDE_PENDING):
        if device is not None and key != (device.index if device.index is not None else torch.cuda.current_device()):
            continue
        torch.cuda.current_stream(torch.device("cuda", key)).wait_stream(_SIDE[key])
        _SIDE_PENDING.discard(key)


def _on_side_stream(fn, tensors):
    """Fork: run ``fn`` on the side stream after everything queued so far on the current stream; the join happens
    at the end of the backward pass (autograd callback) or before the fused update.  In backward, wgrad only feeds
    the optimiser, so its tcgen05 work can overlap the bandwidth-bound BN-backward / dgrad chain of earlier layers."""
    dev = tensors[0].device
    side = _side_stream(dev)
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = fn()
    for x in tensors:
        x.record_stream(side)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_PENDING:
        _SIDE_PENDING.add(key)
        try:
            torch.autograd.Variable._execution_engine.queue_callback(lambda: join_side_stream(dev))
        except Exception:
            pass
    return out


BN_FUSED = True  # single-launch grid-barrier BN kernels (disabled when replicas share one physical GPU)


def bn_workspace(c: int, device) -> torch.Tensor:
    """Zero-initialised workspace (counters + block partials + coefficients) for the single-launch BN kernels."""
    return torch.zeros((_ext().bn_ws_bytes(c) + 3) // 4, dtype=torch.int32, device=device)


def bn_fwd_fused(y, gamma, beta, eps, ws, residual=None, relu=True, running_mean=None, running_var=None, momentum=0.1):
    """Training-mode BN (+residual, +ReLU) in ONE launch: statistics, running-stat update and apply."""
    yh = nhwc(y)
    out = torch.empty_like(yh)
    c = y.shape[1]
    mean = torch.empty(c, dtype=torch.float32, device=y.device)
    rstd = torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_fwd_fused(yh, None if residual is None else nhwc(residual), out, gamma, beta, running_mean, running_var,
                        mean, rstd, ws, eps, momentum, relu)
    return from_nhwc(out), mean, rstd


# Both fusions below are implemented and tested but measured SLOWER inside the step on B200 (gpurun_out/r2ab, ResNet-34
# batch 96: 19.84k img/s with neither, 19.45k with the fused stem tail, 19.56k with the halo dz epilogue) and stay off:
# the fused stem tail repeats the issue-bound arg-max gather in both BN-backward passes, and the halo dgrad of the
# 64-channel stage is bound by its N=64 MMAs' shared-memory reads, which the longer epilogue competes with.
POOL_BWD_FUSED = os.environ.get("FLUXDIST_POOL_BWD_FUSED", "0") == "1"
# dz epilogue in the halo dgrad (64-channel stage).  Lost 0.1-1.4 % while every launch was stream-ordered; with the
# tensor-core kernels chained by programmatic dependent launch it replaces six grid-barrier bn_bwd_fused launches (which
# cannot be PDL dependents) by single-pass bn_bwd_apply: +1.8 % (21,838 -> 22,223 img/s, profiles/r2/policy_ab_pdl.txt)
HALO_DZ = os.environ.get("FLUXDIST_HALO_DZ", "1") == "1"


def bn_bwd_pool_fused(idx, g, y, mean, rstd, gamma, beta, ws, dgamma=None, dbeta=None):
    """Backward of ``bn_relu_pool`` in one launch: ``g`` is the gradient of the POOLED activation ([N,C,P,Q]), ``idx`` the
    arg-max bytes of the forward; returns (dy w.r.t. the raw conv output ``y``, dgamma, dbeta)."""
    yh = nhwc(y)
    c = y.shape[1]
    dy = torch.empty_like(yh)
    dgamma = dgamma if dgamma is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    dbeta = dbeta if dbeta is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_bwd_pool_fused(idx, nhwc(g), yh, mean, rstd, gamma, beta, dgamma, dbeta, dy, ws)
    return from_nhwc(dy), dgamma, dbeta


def bn_bwd_fused(dout, y, out, mean, rstd, gamma, beta, ws, relu=True, want_dres=False, dgamma=None, dbeta=None):
    yh = nhwc(y)
    c = y.shape[1]
    dy = torch.empty_like(yh)
    dres = torch.empty_like(yh) if want_dres else None
    dgamma = dgamma if dgamma is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    dbeta = dbeta if dbeta is not None else torch.empty(c, dtype=torch.float32, device=y.device)
    _ext().bn_bwd_fused(nhwc(dout), yh, nhwc(out) if (relu and out is not None) else None, mean, rstd, gamma, beta,
                        dgamma, dbeta, dy, dres, ws, relu)
    return from_nhwc(dy), dgamma, dbeta, (from_nhwc(dres) if want_dres else None)


def normalize_u8(x_u8: torch.Tensor, cpad: int = 3, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """uint8 NHWC [N,H,W,3] -> bf16 logical NCHW (channels_last) with ``cpad`` channels."""
    n, h, w, _ = x_u8.shape
    out = torch.empty(n, h, w, cpad, dtype=BF16, device=x_u8.device)
    _ext().normalize_u8(x_u8.contiguous(), out, list(mean), list(std))
    return from_nhwc(out)


# --------------------------------------------------------------------------- autograd: loss
class _SoftmaxCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        lg = logits if logits.stride(-1) == 1 else logits.contiguous()
        b = lg.shape[0]
        dl = torch.empty_like(lg)
        row = torch.empty(b, dtype=torch.float32, device=lg.device)
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        ctr = torch.zeros(1, dtype=torch.int32, device=lg.device)
        _ext().softmax_ce(lg, labels.long(), dl, row, loss, ctr, 1.0 / b)
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g.to(dl.dtype), None


def softmax_ce(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Mean softmax cross-entropy with integer labels: loss and dlogits from one kernel (K9)."""
    from . import native_available

    if not (logits.is_cuda and native_available() and logits.dtype in (BF16, torch.float32)):
        return F.cross_entropy(logits.float(), labels.long())
    return _SoftmaxCE.apply(logits, labels)


# --------------------------------------------------------------------------- autograd: conv + BN + act
def _tc_ok(cin: int, cout: int) -> bool:
    return cin % 64 == 0 and cout % 64 == 0


def _is_stem(w_shape, stride, pad, x_shape) -> bool:
    ok = tuple(w_shape) == (64, 3, 7, 7) and stride == 2 and pad == 3
    if ok and x_shape is not None:
        ok = x_shape[2] % 16 == 0 and x_shape[3] % 16 == 0 and x_shape[2] >= 32
    return ok


def _halo_fprop_ok(cred: int, cout: int, hw: int) -> bool:
    """Per-shape kernel choice measured by scripts/conv_bench.py (profiles/conv_bench_r1.json): the halo kernel
    wins when the filter slice is small enough to stay resident in shared memory and the image is large
    (ResNet stage 1); the im2col-TMA kernel wins for the wide, small-image stages."""
    return cred == 64 and cout == 64 and hw >= 28


WGRAD_POLICY = os.environ.get("FLUXDIST_WGRAD", "auto")  # auto | halo | igemm


def _halo_wgrad_ok(cin: int, cout: int, hw: int) -> bool:
    """Weight-gradient kernel choice (scripts/conv_bench_r2.py): the im2col kernel with 128-wide (tap, c) tiles and a
    one-wave split-K wins wherever the input has >= 128 channels; the halo kernel keeps the 64-channel stage, where an
    im2col tile is half empty (64 output channels in a 128-row tile)."""
    if WGRAD_POLICY == "halo":
        return hw >= 14
    if WGRAD_POLICY == "igemm":
        return False
    return cin == 64 and cout == 64 and hw >= 14


class _ConvBNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, gamma, beta, rm, rv, residual, relu, stride, pad, eps, momentum, training, w_lowp,
                w_gradbuf, bn_ws=None, add_link=None, res_link=None, park_link=None, pool=False, own_link=None,
                src_link=None):
        # own_link / src_link: a dict shared by this layer (own_link) and the single conv that consumes its output
        # (src_link there).  The consumer's stride-1 dgrad applies this layer's ReLU mask and reduces the BN-backward
        # sums in its epilogue (conv_dgrad(bn_epi=...)); this layer's backward is then one elementwise pass.
        ctx.own_link, ctx.src_link = own_link, src_link
        # pool: fuse a 3x3/2/pad-1 max-pool behind BN+ReLU (ResNet stem): see bn_relu_pool
        # park_link: given to a 1x1/stride-2 projection shortcut whose input is also read by a stride-2 3x3 conv that
        # got the same dict as add_link.  The projection parks its data gradient as a COMPACT [N,H/2,W/2,C] tensor (a
        # plain GEMM) and the 3x3's parity-decomposed dgrad sums it into its (even, even) class.
        ctx.park_link = park_link
        # add_link / res_link: a dict shared by the two convs of a residual block.  The conv that consumes the
        # skip input (`res_link`) parks its residual gradient there instead of returning it, and the conv that
        # reads the block input (`add_link`) sums it into dx inside the dgrad epilogue — one tensor add per block
        # that autograd would otherwise launch as a separate elementwise kernel.
        ctx.add_link, ctx.res_link = add_link, res_link
        wl = w_lowp if w_lowp is not None else w.detach().to(BF16)
        if not wl.permute(0, 2, 3, 1).is_contiguous():
            wl = wl.contiguous(memory_format=torch.channels_last)
        x = x if x.dtype == BF16 else x.to(BF16)
        tc = _tc_ok(w.shape[1], w.shape[0])
        k3 = tc and w.shape[2] == 3 and w.shape[3] == 3 and stride == 1 and pad == 1
        # BN statistics reduced in the conv epilogue (from the fp32 accumulators): the BN forward is then one
        # elementwise pass without a grid barrier.  The buffers live in the layer workspace and are re-zeroed by the
        # apply kernel itself (its last CTA), so the steady state has no memset either.
        st = None
        if training and EPI_STATS and bn_ws is not None:
            c_ = w.shape[0]
            wf = bn_ws.view(torch.float32)
            st = (wf[4 + 4 * c_: 4 + 5 * c_], wf[4 + 5 * c_: 4 + 6 * c_])
        if k3 and _halo_fprop_ok(w.shape[1], w.shape[0], x.shape[2]):
            y = conv3x3_fprop(x, wl, stats=st if st is not None else False)  # halo-tile kernel: one patch load serves all 9 taps
            y = y[0] if st is not None else y
        elif tc:
            y = conv_fprop(x, wl, stride, pad, stats=st)
            y = y[0] if st is not None else y
        elif _is_stem(w.shape, stride, pad, x.shape):
            orig_shape = tuple(x.shape)
            x2 = stem_s2d(x)  # space-to-depth: the 7x7/2 conv becomes a 4x4/1 conv with K = 16 taps x 16 channels
            y = stem_fprop(x2, wl, stats=st if st is not None else False)
            y = y[0] if st is not None else y
            x = x2  # saved for the weight gradient instead of the image
            ctx.orig_shape = orig_shape
        else:  # other non-64-multiple channel counts fall back to the library
            y = F.conv2d(x, wl, None, stride, pad)
            y = y.contiguous(memory_format=torch.channels_last)
            st = None
        fused_bn = training and BN_FUSED and bn_ws is not None
        ctx.pool_idx = None
        if pool:
            assert relu and residual is None, "fused pool needs a plain BN+ReLU layer"
            if st is not None:
                out, ctx.pool_idx, mean, rstd = bn_relu_pool(y, st[0], st[1], gamma, beta, eps, rm, rv, momentum, True,
                                                             consumed=bn_ws[2:3])
            else:
                if training:
                    s, ss = bn_stats(y)
                else:
                    s = ss = gamma  # unused
                out, ctx.pool_idx, mean, rstd = bn_relu_pool(y, s, ss, gamma, beta, eps, rm, rv, momentum, training)
        elif st is not None:
            out, mean, rstd = bn_apply(y, st[0], st[1], gamma, beta, eps, residual, relu, rm, rv, momentum, True,
                                       consumed=bn_ws[2:3])
        elif fused_bn:
            out, mean, rstd = bn_fwd_fused(y, gamma, beta, eps, bn_ws, residual, relu, rm, rv, momentum)
        else:
            if training:
                s, ss = bn_stats(y)
            else:
                s = ss = gamma  # unused
            out, mean, rstd = bn_apply(y, s, ss, gamma, beta, eps, residual, relu, rm, rv, momentum, training)
        ctx.bn_ws = bn_ws if fused_bn else None
        if own_link is not None and fused_bn and DZ_EPILOGUE and relu and not pool \
                and (residual is None or own_link.get("tail")):
            # residual layers ("tail" links, experimental): the ReLU mask comes from the post-residual activation
            own_link["bn"] = (y, out if residual is not None else None, mean, rstd, gamma, beta, bn_ws)
        # residual layers need the saved activation for the ReLU mask; plain layers recompute it from y
        ctx.save_for_backward(x, wl, y, out if (relu and residual is not None) else None, mean, rstd, gamma, beta)
        stem = (not tc) and _is_stem(w.shape, stride, pad, None)
        ctx.cfg = (relu, stride, pad, residual is not None, tc, tuple(w.shape), tuple(x.shape), k3, stem)
        ctx.w_gradbuf, ctx.g_gradbuf, ctx.b_gradbuf = w_gradbuf if isinstance(w_gradbuf, tuple) else (w_gradbuf, None, None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wl, y, out, mean, rstd, gamma, beta = ctx.saved_tensors
        relu, stride, pad, has_res, tc, w_shape, x_shape, k3, stem = ctx.cfg
        dout = dout if dout.dtype == BF16 else dout.to(BF16)
        pool_fused = ctx.pool_idx is not None and ctx.bn_ws is not None and POOL_BWD_FUSED and relu and not has_res
        if ctx.pool_idx is not None and not pool_fused:  # route the pooled gradient back to full resolution (gather)
            n_, c_, h_, w_ = y.shape
            dfull = torch.empty(n_, h_, w_, c_, dtype=BF16, device=y.device)
            _ext().maxpool_bwd_idx(ctx.pool_idx, nhwc(dout), dfull, 3, 2, 1)
            dout = from_nhwc(dfull)
        if pool_fused:
            # stem tail: max-pool backward + ReLU mask + BN backward in ONE launch; the full-resolution gradient of
            # the pooled activation is never written (bn_bwd_pool_fused_kernel)
            dy, dgamma, dbeta = bn_bwd_pool_fused(ctx.pool_idx, dout, y, mean, rstd, gamma, beta, ctx.bn_ws,
                                                  ctx.g_gradbuf, ctx.b_gradbuf)
            dres = None
        elif ctx.own_link is not None and ctx.own_link.pop("dz", False):
            # dout is already the masked gradient dz; its sums sit in the workspace (reduced by the consumer's dgrad)
            c_ = y.shape[1]
            wf = ctx.bn_ws.view(torch.float32)
            dy, dgamma, dbeta = bn_bwd_apply(dout, y, mean, rstd, gamma, wf[4 + 6 * c_: 4 + 7 * c_],
                                             wf[4 + 7 * c_: 4 + 8 * c_], ctx.g_gradbuf, ctx.b_gradbuf,
                                             consumed=ctx.bn_ws[3:4])
            dres = dout if has_res else None  # the masked gradient is also what flows into the skip connection
        elif ctx.bn_ws is not None:
            dy, dgamma, dbeta, dres = bn_bwd_fused(dout, y, out, mean, rstd, gamma, beta, ctx.bn_ws, relu,
                                                   want_dres=has_res, dgamma=ctx.g_gradbuf, dbeta=ctx.b_gradbuf)
        else:
            dy, dgamma, dbeta, dres = bn_bwd(dout, y, out, mean, rstd, gamma, relu, want_dres=has_res,
                                             dgamma=ctx.g_gradbuf, dbeta=ctx.b_gradbuf, beta=beta)
        if ctx.g_gradbuf is not None:  # landed in the arena: autograd must not accumulate them again
            dgamma = dbeta = None
        if has_res and ctx.res_link is not None:
            ctx.res_link["dres"] = dres  # picked up by the block's first conv (runs later in the backward pass)
            dres = None
        dx = None
        if ctx.needs_input_grad[0]:
            add = ctx.add_link.pop("dres", None) if ctx.add_link is not None else None
            addc = ctx.add_link.pop("dcompact", None) if ctx.add_link is not None else None
            if ctx.add_link is not None and ctx.add_link.get("expect_compact") and addc is None:
                ctx.add_link["late"] = True  # the projection has not run yet: it must return its own dx
            park = ctx.park_link
            if k3 and _halo_fprop_ok(w_shape[0], w_shape[1], x_shape[2]):
                epi = None
                src = ctx.src_link
                if HALO_DZ and src is not None and "bn" in src and addc is None:
                    y_s, out_s, mean_s, rstd_s, gamma_s, beta_s, ws_s = src["bn"]
                    cs = y_s.shape[1]
                    wfs = ws_s.view(torch.float32)
                    epi = (y_s, out_s, mean_s, rstd_s, gamma_s, beta_s, wfs[4 + 6 * cs: 4 + 7 * cs], wfs[4 + 7 * cs: 4 + 8 * cs])
                    src["dz"] = True
                dx = conv3x3_dgrad(dy, wl, add=add, bn_epi=epi)
            elif tc and park is not None and not park.get("late") and stride == 2 and w_shape[2] == 1 \
                    and _s2_native_ok(x_shape[2], x_shape[3], 1, 1, pad):
                dyh = nhwc(dy)
                park["dcompact"] = from_nhwc(
                    gemm_nn(dyh.reshape(-1, dyh.shape[3]), krsc(wl).reshape(w_shape[0], w_shape[1]))
                    .view(dyh.shape[0], dyh.shape[1], dyh.shape[2], w_shape[1]))
            elif tc:
                epi = None
                src = ctx.src_link
                # measured (scripts/dz_bench.py, B200, batch 96): the epilogue work is hidden behind the MMAs when a CTA owns
                # <= 2 output tiles (14x14: dgrad +6 us, BN backward -13 us; 7x7: +5 / -12) but becomes the bottleneck of
                # the 28x28 layers (4 tiles per CTA: +19 / -15), so those keep the two-pass BN backward
                rows = x_shape[0] * x_shape[2] * x_shape[3]
                if src is not None and "bn" in src and stride == 1 and addc is None and rows <= DZ_MAX_ROWS:
                    y_s, out_s, mean_s, rstd_s, gamma_s, beta_s, ws_s = src["bn"]
                    cs = y_s.shape[1]
                    wfs = ws_s.view(torch.float32)
                    epi = (y_s, out_s, mean_s, rstd_s, gamma_s, beta_s, wfs[4 + 6 * cs: 4 + 7 * cs], wfs[4 + 7 * cs: 4 + 8 * cs])
                    src["dz"] = True
                dx = conv_dgrad(dy, wl, x_shape, stride, pad, add=add, add_compact=addc, bn_epi=epi)
            elif stem:
                dx = torch.nn.grad.conv2d_input(ctx.orig_shape, wl, dy, stride, pad)
            else:
                dx = torch.nn.grad.conv2d_input(x_shape, wl, dy, stride, pad)
                if add is not None:
                    dx = dx + add
        side = SIDE_WGRAD and ctx.w_gradbuf is not None and dy.is_cuda
        if stem:
            fn = lambda: stem_wgrad(dy, x, out=ctx.w_gradbuf)  # x holds the s2d tensor here
            gw = _on_side_stream(fn, (dy, x)) if side else fn()
        elif k3 and _halo_wgrad_ok(w_shape[1], w_shape[0], x_shape[2]):
            fn = lambda: conv3x3_wgrad(dy, x, out=ctx.w_gradbuf)
            gw = _on_side_stream(fn, (dy, x)) if side else fn()
        elif tc:
            fn = lambda: conv_wgrad(dy, x, w_shape, stride, pad, out=ctx.w_gradbuf)
            gw = _on_side_stream(fn, (dy, x)) if side else fn()
        else:
            gw = torch.nn.grad.conv2d_weight(x.float(), w_shape, dy.float(), stride, pad)
            if ctx.w_gradbuf is not None:
                ctx.w_gradbuf.add_(gw)
                gw = ctx.w_gradbuf
        # when the gradient already landed in the arena, autograd must not add it a second time
        dw = None if ctx.w_gradbuf is not None else gw
        return (dx, dw, dgamma, dbeta, None, None, dres if has_res else None, None, None, None, None, None, None,
                None, None, None, None, None, None, None, None, None)


def conv_bn_act(x, w, gamma, beta, rm, rv, residual, relu, stride, pad, eps, momentum, training, w_lowp=None,
                w_gradbuf=None, bn_ws=None, add_link=None, res_link=None, park_link=None, pool=False, own_link=None,
                src_link=None):
    return _ConvBNAct.apply(x, w, gamma, beta, rm, rv, residual, relu, stride, pad, eps, momentum, training, w_lowp,
                            w_gradbuf, bn_ws, add_link, res_link, park_link, pool, own_link, src_link)


# --------------------------------------------------------------------------- autograd: max-pool, GAP + FC
class _MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        xh = nhwc(x)
        n, h, w, c = xh.shape
        p, q = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        out = torch.empty(n, p, q, c, dtype=x.dtype, device=x.device)
        idx = torch.empty(n, p, q, c, dtype=torch.uint8, device=x.device)  # window-relative arg-max, 1 byte/element
        _ext().maxpool_fwd(xh, out, idx, k, stride, pad)
        ctx.save_for_backward(idx)
        ctx.cfg = (k, stride, pad, tuple(xh.shape))
        return from_nhwc(out)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        k, stride, pad, xshape = ctx.cfg
        dx = torch.empty(xshape, dtype=BF16, device=idx.device)
        _ext().maxpool_bwd_idx(idx, nhwc(dout.to(BF16)), dx, k, stride, pad)  # gather: no atomics, no memset
        return from_nhwc(dx), None, None, None


def maxpool2d(x, k=3, stride=2, pad=1):
    return _MaxPool.apply(x, k, stride, pad)


class _GapLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, w_lowp, w_gradbuf, b_gradbuf):
        xh = nhwc(x)
        n, h, ww, c = xh.shape
        pooled = torch.empty(n, c, dtype=BF16, device=x.device)
        _ext().gap_fwd(xh, pooled)
        wl = w_lowp if w_lowp is not None else w.detach().to(BF16)
        logits = gemm_nt(pooled, wl)                       # tcgen05, [n, classes]
        if b is not None:
            _ext().bias_add(logits, b.detach().float())
        ctx.save_for_backward(pooled, wl)
        ctx.cfg = (tuple(xh.shape), b is not None)
        ctx.bufs = (w_gradbuf, b_gradbuf)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        pooled, wl = ctx.saved_tensors
        (n, h, ww, c), has_b = ctx.cfg
        w_gradbuf, b_gradbuf = ctx.bufs
        d = dlogits.contiguous().to(BF16)
        ncls = d.shape[1]
        e = _ext()
        # dW[cls, c] = sum_n d[n, cls] * pooled[n, c]
        gw = w_gradbuf if w_gradbuf is not None else torch.empty(ncls, c, dtype=torch.float32, device=d.device)
        tc = c % 128 == 0 and ncls % 8 == 0
        if tc:
            e.gemm_tn(d, pooled, gw)        # tcgen05: A = d stored [K=n, M=cls], B = pooled [K=n, N=c]
        else:
            e.small_gemm(d, pooled, gw, ncls, c, n, 1, ncls, c, 1)
        gb = None
        if has_b:
            gb = b_gradbuf if b_gradbuf is not None else torch.empty(ncls, dtype=torch.float32, device=d.device)
            e.colsum(d, gb)
        # dpooled[n, c] = sum_cls d[n, cls] * W[cls, c]
        dp = torch.empty(n, c, dtype=BF16, device=d.device)
        if tc:
            e.gemm_nn(d, wl.contiguous(), dp)  # tcgen05: d[n, cls] @ W[cls, c]
        else:
            e.small_gemm(d, wl, dp, n, c, ncls, ncls, 1, c, 1)
        dx = torch.empty(n, h, ww, c, dtype=BF16, device=d.device)
        e.gap_bwd(dp, dx)
        return (from_nhwc(dx), None if w_gradbuf is not None else gw,
                (None if (b_gradbuf is not None or not has_b) else gb), None, None, None)


def gap_linear(x, w, b, w_lowp=None):
    return _GapLinear.apply(x, w, b, w_lowp, getattr(w, "_gradbuf", None),
                            getattr(b, "_gradbuf", None) if b is not None else None)


# --------------------------------------------------------------------------- autograd: plain Conv2d / Linear (any channel count)
class _DirectConv(torch.autograd.Function):
    """``Conv2d`` (with bias) on the direct CUDA-core kernels (csrc/direct_conv.cu): the path for channel counts the
    tensor-core tiles do not cover — the reference's test model ``Conv((7,7),3=>3)`` (test/single_device.jl:117-119)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, w_gradbuf, b_gradbuf):
        xh = nhwc(x if x.dtype == BF16 else x.to(BF16))
        wk = w.detach().to(BF16).permute(0, 2, 3, 1).contiguous()  # [K,R,S,C]
        n, h, ww, c = xh.shape
        k, r, s_, _ = wk.shape
        p, q = (h + 2 * pad - r) // stride + 1, (ww + 2 * pad - s_) // stride + 1
        y = torch.empty(n, p, q, k, dtype=BF16, device=x.device)
        _ext().direct_conv_fprop(xh, wk, None if b is None else b.detach().float(), y, stride, pad)
        ctx.save_for_backward(xh, wk)
        ctx.cfg = (stride, pad, tuple(w.shape), b is not None)
        ctx.bufs = (w_gradbuf, b_gradbuf)
        return from_nhwc(y)

    @staticmethod
    def backward(ctx, dy):
        xh, wk = ctx.saved_tensors
        stride, pad, w_shape, has_b = ctx.cfg
        w_gradbuf, b_gradbuf = ctx.bufs
        dyh = nhwc(dy if dy.dtype == BF16 else dy.to(BF16))
        e = _ext()
        dx = None
        if ctx.needs_input_grad[0]:
            dxh = torch.empty_like(xh)
            e.direct_conv_dgrad(dyh, wk, dxh, stride, pad)
            dx = from_nhwc(dxh)
        k, c, r, s_ = w_shape
        gw = w_gradbuf if w_gradbuf is not None else torch.zeros(w_shape, dtype=torch.float32, device=dy.device)
        gb = None
        if has_b:
            gb = b_gradbuf if b_gradbuf is not None else torch.zeros(k, dtype=torch.float32, device=dy.device)
        st = gw.stride()  # element strides of the logical [K,C,R,S] gradient buffer, whatever its memory format
        e.direct_conv_wgrad(dyh, xh, gw, gb, r, s_, stride, pad, st[0], st[2], st[3], st[1])
        return (dx, None if w_gradbuf is not None else gw, None if (b_gradbuf is not None or not has_b) else gb,
                None, None, None, None)


def direct_conv2d(x, w, b, stride=1, pad=0):
    return _DirectConv.apply(x, w, b, stride, pad, getattr(w, "_gradbuf", None),
                             getattr(b, "_gradbuf", None) if b is not None else None)


class _Linear(torch.autograd.Function):
    """``Linear`` on this repository's GEMM kernels: tcgen05 when the reduction dimension is a multiple of 64, the
    small-GEMM kernel otherwise (e.g. Dense(2028, 10) of the reference's test model)."""

    @staticmethod
    def forward(ctx, x, w, b, w_lowp, w_gradbuf, b_gradbuf):
        x2 = (x if x.dtype == BF16 else x.to(BF16)).contiguous()
        wl = (w_lowp if w_lowp is not None else w.detach().to(BF16)).contiguous()
        n, kdim = x2.shape
        m = wl.shape[0]
        e = _ext()
        if kdim % 64 == 0:
            out = gemm_nt(x2, wl)
        else:
            out = torch.empty(n, m, dtype=BF16, device=x.device)
            e.small_gemm(x2, wl, out, n, m, kdim, kdim, 1, 1, kdim)  # out[n,m] = sum_k x[n,k] * w[m,k]
        if b is not None:
            e.bias_add(out, b.detach().float())
        ctx.save_for_backward(x2, wl)
        ctx.has_b = b is not None
        ctx.bufs = (w_gradbuf, b_gradbuf)
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, wl = ctx.saved_tensors
        w_gradbuf, b_gradbuf = ctx.bufs
        d = dout.contiguous().to(BF16)
        n, kdim = x2.shape
        m = wl.shape[0]
        e = _ext()
        gw = torch.zeros(m, kdim, dtype=torch.float32, device=d.device)
        e.small_gemm(d, x2, gw, m, kdim, n, 1, m, kdim, 1)      # gw[m,k] = sum_n d[n,m] * x[n,k]
        if w_gradbuf is not None:
            w_gradbuf.add_(gw)
        gb = None
        if ctx.has_b:
            gb = torch.empty(m, dtype=torch.float32, device=d.device)
            e.colsum(d, gb)
            if b_gradbuf is not None:
                b_gradbuf.add_(gb)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(n, kdim, dtype=BF16, device=d.device)
            e.small_gemm(d, wl, dx, n, kdim, m, m, 1, kdim, 1)  # dx[n,k] = sum_m d[n,m] * w[m,k]
        return (dx, None if w_gradbuf is not None else gw, None if (b_gradbuf is not None or not ctx.has_b) else gb,
                None, None, None)


def linear(x, w, b, w_lowp=None):
    return _Linear.apply(x, w, b, w_lowp, getattr(w, "_gradbuf", None),
                         getattr(b, "_gradbuf", None) if b is not None else None)


def fused_update_local(*a, **k):  # kept for API symmetry; the engine builds FusedPlan objects directly
    raise NotImplementedError("use parallel.native_comm.NativeComm")
