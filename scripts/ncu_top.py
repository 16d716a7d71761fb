"""One launch (after one warm launch) of each hot kernel on a representative ResNet-34 batch-96 shape, for
`ncu --set full` (scripts/ncu_top.sh).  Not a benchmark: numbers under the profiler are never reported as results."""
import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
from fluxdist.ops import native as N
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
B = 96
def conv_set(hw, C, K):
    x = cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)); dy = cl(torch.randn(B, K, hw, hw, device="cuda").to(BF))
    w = cl((torch.randn(K, C, 3, 3, device="cuda") * 0.05).to(BF))
    gb = torch.zeros(K, C, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last)
    return x, dy, w, gb
for rep in range(2):
    x, dy, w, gb = conv_set(28, 128, 128)
    ops.conv_fprop(x, w, 1, 1)                       # igemm FPROP (cluster + multicast)
    ops.conv_dgrad(dy, w, x.shape, 1, 1)             # igemm DGRAD
    ops.conv3x3_wgrad(dy, x, out=gb)                 # halo wgrad
    x, dy, w, gb = conv_set(56, 64, 64)
    ops.conv3x3_fprop(x, w, stats=True)              # halo fprop + BN statistics
    ops.conv3x3_dgrad(dy, w)                         # halo dgrad
    ops.conv3x3_wgrad(dy, x, out=gb)
    x, dy, w, gb = conv_set(7, 512, 512)
    ops.conv_wgrad(dy, x, w.shape, 1, 1, out=gb)     # igemm WGRAD (split-K)
    for hw, C in [(56, 64), (14, 256)]:
        y = cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)); d = cl(torch.randn(B, C, hw, hw, device="cuda").to(BF))
        gamma = torch.rand(C, device="cuda") + 0.5; beta = torch.randn(C, device="cuda")
        ws = N.bn_workspace(C, "cuda"); rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        out, mean, rstd = N.bn_fwd_fused(y, gamma, beta, 1e-5, ws, d, True, rm, rv)
        N.bn_bwd_fused(d, y, out, mean, rstd, gamma, beta, ws, True, want_dres=True)
    torch.cuda.synchronize()
