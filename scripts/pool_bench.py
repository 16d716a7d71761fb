"""Stem tail kernels at ResNet-34 batch-96 shape: fused BN+ReLU+pool forward, arg-max gather backward (graph-timed)."""
import json, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
from fluxdist.ops import native as N
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
NSET, NL, B = 3, 9, 96
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
ys = [cl(torch.randn(B, 64, 112, 112, device="cuda").to(BF)) for _ in range(NSET)]
gamma, beta = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda")
s, ss = ops.bn_stats(ys[0])
pooled, idx, _, _ = N.bn_relu_pool(ys[0], s, ss, gamma, beta, 1e-5)
dfs = [torch.empty(B, 112, 112, 64, dtype=BF, device="cuda") for _ in range(NSET)]
gs = [torch.randn(B, 56, 56, 64, device="cuda").to(BF) for _ in range(NSET)]
r = {}
r["bn_relu_pool_us"] = timeit(lambda i: (lambda: N.bn_relu_pool(ys[i], s, ss, gamma, beta, 1e-5)))
r["maxpool_bwd_idx_us"] = timeit(lambda i: (lambda: ops.extension().maxpool_bwd_idx(idx, gs[i], dfs[i], 3, 2, 1)))
r["maxpool_fwd_us"] = timeit(lambda i: (lambda: ops.maxpool2d(ys[i], 3, 2, 1)))
print(json.dumps(r))
