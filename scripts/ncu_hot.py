"""One launch (after one warm repetition) of every hot kernel on its ResNet-34 batch-96 shape, for `ncu --set full`
(scripts/ncu_hot.sh).  Not a benchmark: numbers under the profiler are never reported as results.
Kernels matched by the ncu regex, in order (17 per repetition): see LABELS."""
import sys, torch
sys.path.insert(0, ".")
import fluxdist as fd
from fluxdist import ops
from fluxdist.ops import native as N
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
B = 96
LABELS = ["igemm FPROP+stats 3x3 28x28x128 (256x128 tile, 8 epilogue warps)", "igemm FPROP+stats 3x3 14x14x256 (256x128 tile)",
          "igemm DGRAD 3x3 28x28x128 (256x128 tile)", "igemm DGRAD+dz epilogue 3x3 14x14x256 (128x128 tiles)",
          "igemm DGRAD_S2 (parity) 3x3 28x28x128 -> 56x56x64", "igemm WGRAD 3x3 14x14x256 (128x128 tiles, split-K 4)",
          "igemm WGRAD 3x3 7x7x512 (128x128 tiles, no split)", "halo FPROP+stats 3x3 56x56x64", "halo DGRAD 3x3 56x56x64",
          "halo WGRAD 3x3 56x56x64", "bn_apply 28x28x128 + residual", "bn_bwd_apply 14x14x256", "bn_bwd_fused 56x56x64 + residual",
          "bn_relu_pool 112x112x64", "maxpool_bwd_idx 112x112x64", "stem FPROP (7x7/2 as 4x4/1 over s2d) 224x224x3",
          "fused_allreduce_update (1 GPU: update only) 21.8M parameters"]
def conv_set(hw, C, K, R=3, ho=None):
    ho = ho or hw
    x = cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)); dy = cl(torch.randn(B, K, ho, ho, device="cuda").to(BF))
    w = cl((torch.randn(K, C, R, R, device="cuda") * 0.05).to(BF))
    gb = torch.zeros(K, C, R, R, device="cuda").contiguous(memory_format=torch.channels_last)
    return x, dy, w, gb
def bn_set(C):
    return torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda"), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
m34 = fd.ResNet34(1000).cuda().to(memory_format=torch.channels_last)
from fluxdist.parallel import FusedGroup
grp = FusedGroup.for_threads([m34], fd.cuda_devices(1), fd.Momentum(0.01, 0.9), comm="native")
rep = grp.replicas[0]
for r in range(2):
    x, dy, w, gb = conv_set(28, 128, 128)
    sm, sq = torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda")
    ops.conv_fprop(x, w, 1, 1, stats=(sm, sq))                                            # 1
    x2, dy2, w2, gb2 = conv_set(14, 256, 256)
    s2, q2 = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    ops.conv_fprop(x2, w2, 1, 1, stats=(s2, q2))                                          # 2
    ops.conv_dgrad(dy, w, x.shape, 1, 1)                                                  # 3
    g2, b2, _, _ = bn_set(256)
    yy = cl(torch.randn(B, 256, 14, 14, device="cuda").to(BF)); m2 = torch.zeros(256, device="cuda"); r2 = torch.ones(256, device="cuda")
    s_, ss_ = torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda")
    dz = ops.conv_dgrad(dy2, w2, x2.shape, 1, 1, bn_epi=(yy, None, m2, r2, g2, b2, s_, ss_))  # 4
    x3, dy3, w3, _ = conv_set(56, 64, 128, 3, 28)
    ops.conv_dgrad(dy3, w3, x3.shape, 2, 1)                                               # 5
    ops.conv_wgrad(dy2, x2, w2.shape, 1, 1, out=gb2)                                      # 6
    x7, dy7, w7, gb7 = conv_set(7, 512, 512)
    ops.conv_wgrad(dy7, x7, w7.shape, 1, 1, out=gb7)                                      # 7
    x5, dy5, w5, gb5 = conv_set(56, 64, 64)
    s5, q5 = torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda")
    ops.conv3x3_fprop(x5, w5, stats=(s5, q5))                                             # 8
    ops.conv3x3_dgrad(dy5, w5)                                                            # 9
    ops.conv3x3_wgrad(dy5, x5, out=gb5)                                                   # 10
    gamma, beta, rm, rv = bn_set(128)
    y28 = cl(torch.randn(B, 128, 28, 28, device="cuda").to(BF)); res = cl(torch.randn(B, 128, 28, 28, device="cuda").to(BF))
    sA, qA = ops.bn_stats(y28)   # (bn_stats_kernel is not in the regex)
    N.bn_apply(y28, sA, qA, gamma, beta, 1e-5, res, True, rm, rv)                         # 11
    N.bn_bwd_apply(dz, yy, m2, r2, g2, s_, ss_)                                           # 12
    y56 = cl(torch.randn(B, 64, 56, 56, device="cuda").to(BF)); d56 = cl(torch.randn(B, 64, 56, 56, device="cuda").to(BF))
    g6, b6, _, _ = bn_set(64)
    ws = N.bn_workspace(64, "cuda")
    N.bn_bwd_fused(d56, y56, y56, torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), g6, b6, ws, True, want_dres=True)  # 13
    y112 = cl(torch.randn(B, 64, 112, 112, device="cuda").to(BF))
    s, ss = torch.zeros(64, device="cuda"), torch.full((64,), float(B * 112 * 112), device="cuda")
    pooled, idx, _, _ = N.bn_relu_pool(y112, s, ss, g6, b6, 1e-5)                          # 14
    dfull = torch.empty(B, 112, 112, 64, dtype=BF, device="cuda")
    ops.extension().maxpool_bwd_idx(idx, pooled.permute(0, 2, 3, 1), dfull, 3, 2, 1)       # 15
    img = cl(torch.randn(B, 3, 224, 224, device="cuda").to(BF))
    x2s = N.stem_s2d(img)
    wst = cl((torch.randn(64, 3, 7, 7, device="cuda") * 0.05).to(BF))
    N.stem_fprop(x2s, wst, stats=(s5, q5))                                                # 16
    rep.arena.g32.normal_()
    rep.reduce_and_update()                                                               # 17
    torch.cuda.synchronize()
print("\n".join(f"{i + 1:2d} {l}" for i, l in enumerate(LABELS)))
