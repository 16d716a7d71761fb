"""Cost of the BN-statistics epilogue in the im2col fprop kernel, per ResNet-34 shape (graph-timed, rotating buffers)."""
import json, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
NSET, NL, B = 4, 12, 96
def timeit(make_fn):
    fns = [make_fn(i) for i in range(NSET)]
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(NL): fns[i % NSET]()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / (5 * NL) * 1e3, 1)
for hw, C, K, R, st in [(28, 128, 128, 3, 1), (56, 64, 128, 3, 2), (56, 64, 128, 1, 2), (14, 256, 256, 3, 1), (28, 128, 256, 3, 2), (7, 512, 512, 3, 1), (14, 256, 512, 3, 2)]:
    xs = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    w = cl((torch.randn(K, C, R, R, device="cuda") * 0.05).to(BF))
    sm, sq = torch.zeros(K, device="cuda"), torch.zeros(K, device="cuda")
    r = {"hw": hw, "C": C, "K": K, "R": R, "stride": st}
    r["fprop_us"] = timeit(lambda i: (lambda: ops.conv_fprop(xs[i], w, st, R // 2)))
    r["fprop_stats_us"] = timeit(lambda i: (lambda: ops.conv_fprop(xs[i], w, st, R // 2, stats=(sm, sq))))
    print(json.dumps(r), flush=True)
