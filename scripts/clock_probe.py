import sys, time, subprocess, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
from fluxdist.utils.tracing import ClockSampler
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
def loop(name, fn, secs=2.0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    cs = ClockSampler(0, 50).start()
    t0 = time.time(); n = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.time() - t0 < secs:
        g.replay(); n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    r = cs.stop()
    print(f"{name}: {a.elapsed_time(b)/n*1e3:.1f} us/launch, clocks {r}", flush=True)
x = cl(torch.randn(96, 64, 56, 56, device="cuda").to(BF)); w = cl((torch.randn(64, 64, 3, 3, device="cuda") * 0.05).to(BF))
x2 = cl(torch.randn(96, 128, 28, 28, device="cuda").to(BF)); w2 = cl((torch.randn(128, 128, 3, 3, device="cuda") * 0.05).to(BF))
loop("halo_fprop s1 v1", lambda: ops.conv3x3_fprop(x, w, variant=1))
loop("halo_fprop s2 v3", lambda: ops.conv3x3_fprop(x2, w2, variant=3))
loop("im2col_fprop s2", lambda: ops.conv_fprop(x2, w2, 1, 1))
loop("cudnn_fprop s1", lambda: torch.nn.functional.conv2d(x, w, None, 1, 1))
a = torch.randn(8192, 8192, device="cuda").to(BF); b2 = torch.randn(8192, 8192, device="cuda").to(BF)
loop("cublas 8192^3", lambda: torch.matmul(a, b2))
loop("gemm_nt 8192^3", lambda: ops.gemm_nt(a, b2))
