"""CPU checks of the PyTorch oracles the GPU numerics tests compare the sm_100a kernels against, and of the pure-Python
dispatch policies of the native path (no GPU needed)."""
import torch
import torch.nn.functional as F

import fluxdist  # noqa: F401
from fluxdist import ops
from fluxdist.ops import native as NV


def test_bn_bwd_apply_ref_matches_autograd():
    torch.manual_seed(0)
    y = torch.randn(3, 8, 7, 5)
    g, b = torch.randn(8), torch.randn(8)
    dz = torch.randn_like(y)
    mean = y.mean((0, 2, 3))
    rstd = torch.rsqrt(y.var((0, 2, 3), unbiased=False) + 1e-5)
    dy, dg, db = ops.bn_bwd_apply_ref(dz, y, mean, rstd, g)
    yy = y.clone().requires_grad_(True)
    gg, bb = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.batch_norm(yy, None, None, gg, bb, True, 0.1, 1e-5).backward(dz)
    assert torch.allclose(dy, yy.grad, atol=1e-5)
    assert torch.allclose(dg, gg.grad, atol=1e-4) and torch.allclose(db, bb.grad, atol=1e-4)


def test_bn_relu_pool_ref_shapes_and_monotonicity_argument():
    """The fused kernel pools the RAW conv output (by max where the BN scale is >= 0, by min where it is < 0) and applies
    the affine + ReLU to the pooled value only: check that identity on the oracle."""
    torch.manual_seed(1)
    y = torch.randn(2, 6, 11, 9)
    g = torch.tensor([1.0, -0.7, 0.3, -2.0, 0.0, 1.5])
    b = torch.randn(6) * 0.3
    ref = ops.bn_relu_pool_ref(y, g, b)
    assert ref.shape == (2, 6, 6, 5)
    mean = y.mean((0, 2, 3))
    rstd = torch.rsqrt(y.var((0, 2, 3), unbiased=False) + 1e-5)
    sc = (g * rstd)[None, :, None, None]
    sh = (b - mean * g * rstd)[None, :, None, None]
    pooled_max = F.max_pool2d(y, 3, 2, 1)
    pooled_min = -F.max_pool2d(-y, 3, 2, 1)
    sel = torch.where(sc >= 0, pooled_max, pooled_min)
    assert torch.allclose(torch.relu(sel * sc + sh), ref, atol=1e-5)


def test_stride2_parity_policy():
    assert NV._s2_native_ok(56, 56, 3, 3, 1)        # ResNet stage transitions
    assert NV._s2_native_ok(56, 56, 1, 1, 0)        # projection shortcuts
    assert not NV._s2_native_ok(224, 224, 7, 7, 3)  # filters wider than 3 fall back to zero-upsampling
    assert not NV._s2_native_ok(55, 56, 3, 3, 1)    # odd images
    assert not NV._s2_native_ok(56, 56, 3, 3, 0)    # output is not H/2


def test_stride2_parity_classes_cover_all_taps_once():
    """Mirror of the class enumeration in csrc/igemm_sm100.cu::fd_conv_dgrad_s2: every filter tap lands in exactly one
    (h%2, w%2) class and the classes of a 3x3/pad-1 filter have 4+2+2+1 taps."""
    def classes(R, S, pad):
        out = {}
        for ph in (0, 1):
            for pw in (0, 1):
                out[(ph, pw)] = [(r, s) for r in range(R) for s in range(S)
                                 if (ph + pad - r) % 2 == 0 and (pw + pad - s) % 2 == 0]
        return out
    c = classes(3, 3, 1)
    assert sorted(len(v) for v in c.values()) == [1, 2, 2, 4]
    assert sorted(t for v in c.values() for t in v) == [(r, s) for r in range(3) for s in range(3)]
    c1 = classes(1, 1, 0)
    assert c1[(0, 0)] == [(0, 0)] and all(not v for k, v in c1.items() if k != (0, 0))
    # data-gradient identity on a tiny example: the classes reproduce conv2d_input
    torch.manual_seed(0)
    dy, w = torch.randn(1, 2, 3, 3), torch.randn(2, 3, 3, 3)
    ref = torch.nn.grad.conv2d_input((1, 3, 6, 6), w, dy, 2, 1)
    dx = torch.zeros(1, 3, 6, 6)
    for (ph, pw), taps in c.items():
        for r, s in taps:
            dh, dw = (ph + 1 - r) // 2, (pw + 1 - s) // 2
            for i in range(3):
                for j in range(3):
                    ii, jj = i + dh, j + dw
                    if 0 <= ii < 3 and 0 <= jj < 3:
                        dx[0, :, 2 * i + ph, 2 * j + pw] += torch.einsum("k,kc->c", dy[0, :, ii, jj], w[:, :, r, s])
    assert torch.allclose(dx, ref, atol=1e-5)
