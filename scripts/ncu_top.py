"""One launch (after one warm launch) of each hot kernel on a representative ResNet-34 batch-96 shape, for
`ncu --set full` (scripts/ncu_top.sh).  Not a benchmark: numbers under the profiler are never reported as results.
Kernels matched by the ncu regex, in order (13 per repetition):
  igemm FPROP+stats 28x28x128 | bn_apply | igemm DGRAD+dz epilogue 14x14x256 | bn_bwd_apply | igemm DGRAD_S2 56->28 |
  wgrad_halo 28x28x128 | conv_halo FPROP+stats 56x56x64 | conv_halo DGRAD | wgrad_halo 56x56x64 | igemm WGRAD 7x7x512 |
  bn_bwd_fused 56x56x64 (+residual) | bn_relu_pool 112x112x64 | maxpool_bwd_idx 112x112x64"""
import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
from fluxdist.ops import native as N
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
B = 96
def conv_set(hw, C, K, R=3, ho=None):
    ho = ho or hw
    x = cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)); dy = cl(torch.randn(B, K, ho, ho, device="cuda").to(BF))
    w = cl((torch.randn(K, C, R, R, device="cuda") * 0.05).to(BF))
    gb = torch.zeros(K, C, R, R, device="cuda").contiguous(memory_format=torch.channels_last)
    return x, dy, w, gb
def bn_set(C):
    return torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda"), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
for rep in range(2):
    x, dy, w, gb = conv_set(28, 128, 128)
    gamma, beta, rm, rv = bn_set(128)
    sm, sq = torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda"); cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    y, _, _ = ops.conv_fprop(x, w, 1, 1, stats=(sm, sq))                                   # 1 igemm FPROP + stats
    out, mean, rstd = N.bn_apply(y, sm, sq, gamma, beta, 1e-5, None, True, rm, rv, consumed=cnt)  # 2 bn_apply
    x2, dy2, w2, _ = conv_set(14, 256, 256)
    g2, b2, _, _ = bn_set(256)
    s_, ss_ = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    yy = cl(torch.randn(B, 256, 14, 14, device="cuda").to(BF)); m2 = torch.zeros(256, device="cuda"); r2 = torch.ones(256, device="cuda")
    dz = ops.conv_dgrad(dy2, w2, x2.shape, 1, 1, bn_epi=(yy, None, m2, r2, g2, b2, s_, ss_))  # 3 igemm DGRAD + dz epilogue
    N.bn_bwd_apply(dz, yy, m2, r2, g2, s_, ss_)                                          # 4 bn_bwd_apply
    x3, dy3, w3, _ = conv_set(56, 64, 128, 3, 28)
    ops.conv_dgrad(dy3, w3, x3.shape, 2, 1)                                              # 5 igemm DGRAD_S2
    ops.conv3x3_wgrad(dy, x, out=gb)                                                     # 6 halo wgrad 28
    x, dy, w, gb = conv_set(56, 64, 64)
    ops.conv3x3_fprop(x, w, stats=True)                                                  # 7 halo fprop + stats
    ops.conv3x3_dgrad(dy, w)                                                             # 8 halo dgrad
    ops.conv3x3_wgrad(dy, x, out=gb)                                                     # 9 halo wgrad 56
    x, dy, w, gb = conv_set(7, 512, 512)
    ops.conv_wgrad(dy, x, w.shape, 1, 1, out=gb)                                         # 10 igemm WGRAD
    y = cl(torch.randn(B, 64, 56, 56, device="cuda").to(BF)); d = cl(torch.randn(B, 64, 56, 56, device="cuda").to(BF))
    gamma, beta, rm, rv = bn_set(64)
    ws = N.bn_workspace(64, "cuda")
    s, ss = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
    mean, rstd = torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
    N.bn_bwd_fused(d, y, y, mean, rstd, gamma, beta, ws, True, want_dres=True)          # 11 bn_bwd_fused
    y = cl(torch.randn(B, 64, 112, 112, device="cuda").to(BF))
    s, ss = torch.zeros(64, device="cuda"), torch.full((64,), float(B * 112 * 112), device="cuda")
    pooled, idx, _, _ = N.bn_relu_pool(y, s, ss, gamma, beta, 1e-5)                      # 12 bn_relu_pool
    dfull = torch.empty(B, 112, 112, 64, dtype=BF, device="cuda")
    ops.extension().maxpool_bwd_idx(idx, pooled.permute(0, 2, 3, 1), dfull, 3, 2, 1)     # 13 maxpool_bwd_idx
    torch.cuda.synchronize()
