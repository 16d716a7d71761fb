"""Achieved HBM bandwidth of the BN kernels per ResNet-34 activation shape (graph-timed, rotating buffers)."""
import json, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
NSET, NL = 4, 12
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
    return a.elapsed_time(b) / (5 * NL) * 1e3
for hw, C in [(112, 64), (56, 64), (28, 128), (14, 256), (7, 512)]:
    ys = [cl(torch.randn(96, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    ds = [cl(torch.randn(96, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    gamma = torch.rand(C, device="cuda") + 0.5; beta = torch.randn(C, device="cuda")
    V = ys[0].numel() * 2 / 1e6  # MB
    s, ss = ops.bn_stats(ys[0]); out, mean, rstd = ops.bn_apply(ys[0], s, ss, gamma, beta, 1e-5)
    r = {"hw": hw, "C": C, "MB": round(V, 1)}
    t = timeit(lambda i: (lambda: ops.bn_stats(ys[i]))); r["stats_us"] = round(t, 1); r["stats_GBs"] = round(V / t * 1e3)
    t = timeit(lambda i: (lambda: ops.bn_apply(ys[i], s, ss, gamma, beta, 1e-5))); r["apply_us"] = round(t, 1); r["apply_GBs"] = round(2 * V / t * 1e3)
    t = timeit(lambda i: (lambda: ops.bn_apply(ys[i], s, ss, gamma, beta, 1e-5, residual=ds[i]))); r["apply_res_us"] = round(t, 1); r["apply_res_GBs"] = round(3 * V / t * 1e3)
    t = timeit(lambda i: (lambda: ops.bn_bwd(ds[i], ys[i], None, mean, rstd, gamma, True, beta=beta))); r["bwd_us"] = round(t, 1); r["bwd_GBs"] = round(5 * V / t * 1e3)
    t = timeit(lambda i: (lambda: ops.bn_bwd(ds[i], ys[i], out, mean, rstd, gamma, True, want_dres=True, beta=beta))); r["bwd_res_us"] = round(t, 1); r["bwd_res_GBs"] = round(8 * V / t * 1e3)
    from fluxdist.ops import native as N
    wss = [N.bn_workspace(C, "cuda") for _ in range(NSET)]
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    t = timeit(lambda i: (lambda: N.bn_fwd_fused(ys[i], gamma, beta, 1e-5, wss[i], None, True, rm, rv))); r["F_fwd_us"] = round(t, 1); r["F_fwd_GBs"] = round(3 * V / t * 1e3)
    t = timeit(lambda i: (lambda: N.bn_fwd_fused(ys[i], gamma, beta, 1e-5, wss[i], ds[i], True, rm, rv))); r["F_fwd_res_us"] = round(t, 1)
    t = timeit(lambda i: (lambda: N.bn_bwd_fused(ds[i], ys[i], None, mean, rstd, gamma, beta, wss[i], True))); r["F_bwd_us"] = round(t, 1); r["F_bwd_GBs"] = round(5 * V / t * 1e3)
    t = timeit(lambda i: (lambda: N.bn_bwd_fused(ds[i], ys[i], out, mean, rstd, gamma, beta, wss[i], True, want_dres=True))); r["F_bwd_res_us"] = round(t, 1)
    t = timeit(lambda i: (lambda: N.bn_apply(ys[i], s, ss, gamma, beta, 1e-5, residual=ds[i]))); r["apply_only_res_us"] = round(t, 1)
    t = timeit(lambda i: (lambda: ys[i].add_(ds[i]))); r["torch_add_us"] = round(t, 1); r["torch_add_GBs"] = round(3 * V / t * 1e3)
    print(json.dumps(r), flush=True)
