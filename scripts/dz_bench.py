"""dz epilogue A/B: dgrad (plain vs mask+reduce epilogue) and BN backward (grid-barrier two-pass vs single-pass apply)."""
import json, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
from fluxdist.ops import native as N
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
for hw, C in [(28, 128), (14, 256), (7, 512)]:
    ys = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    ds = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    w = cl((torch.randn(C, C, 3, 3, device="cuda") * 0.05).to(BF))
    gamma = torch.rand(C, device="cuda") + 0.5; beta = torch.randn(C, device="cuda")
    s, ss = ops.bn_stats(ys[0]); _, mean, rstd = ops.bn_apply(ys[0], s, ss, gamma, beta, 1e-5)
    s0, s1 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    wss = [N.bn_workspace(C, "cuda") for _ in range(NSET)]
    r = {"hw": hw, "C": C}
    ext = ops.extension()
    adds = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    for bn_, mt_ in [(0, 0), (128, 1), (128, 2), (256, 1)]:
        if bn_ and C % bn_:
            continue
        ext.set_tile_shape(bn_, mt_)
        tag = "auto" if not bn_ else f"{bn_}x{mt_}"
        r[f"dgrad_{tag}"] = timeit(lambda i: (lambda: ops.conv_dgrad(ds[i], w, ys[i].shape, 1, 1)))
        r[f"dgrad_dz_{tag}"] = timeit(lambda i: (lambda: ops.conv_dgrad(ds[i], w, ys[i].shape, 1, 1, bn_epi=(ys[i], None, mean, rstd, gamma, beta, s0, s1))))
        r[f"dgrad_dz_add_out_{tag}"] = timeit(lambda i: (lambda: ops.conv_dgrad(ds[i], w, ys[i].shape, 1, 1, add=adds[i], bn_epi=(ys[i], adds[(i + 1) % NSET], mean, rstd, gamma, beta, s0, s1))))
    ext.set_tile_shape(0, 0)
    r["bn_bwd_fused_us"] = timeit(lambda i: (lambda: N.bn_bwd_fused(ds[i], ys[i], None, mean, rstd, gamma, beta, wss[i], True)))
    r["bn_bwd_apply_us"] = timeit(lambda i: (lambda: N.bn_bwd_apply(ds[i], ys[i], mean, rstd, gamma, s0, s1)))
    print(json.dumps(r), flush=True)
