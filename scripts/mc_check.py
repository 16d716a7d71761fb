"""Cluster + TMA-multicast variant of the im2col implicit GEMM: correctness against cuDNN and timing on / off.
Same timing method as conv_bench.py (CUDA graph of 12 launches over 4 rotating buffer sets)."""
import json, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
import torch.nn.functional as F
BF = torch.bfloat16
B = 96
ext = ops.extension()
# (hw_in, C, K, R, stride, pad)
shapes = [(28, 128, 128, 3, 1, 1), (56, 64, 128, 3, 2, 1), (56, 64, 128, 1, 2, 0), (14, 256, 256, 3, 1, 1),
          (28, 128, 256, 3, 2, 1), (7, 512, 512, 3, 1, 1), (14, 256, 512, 3, 2, 1)]
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
    return round(a.elapsed_time(b) / (5 * NL) * 1e3, 1)
rows = []
for hw, C, K, R, st, pad in shapes:
    ho = (hw + 2 * pad - R) // st + 1
    xs = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    dys = [cl(torch.randn(B, K, ho, ho, device="cuda").to(BF)) for _ in range(NSET)]
    w = cl((torch.randn(K, C, R, R, device="cuda") * 0.05).to(BF))
    r = {"hw": hw, "C": C, "K": K, "R": R, "stride": st}
    yref = F.conv2d(xs[0].float(), w.float(), None, st, pad)
    dref = torch.nn.grad.conv2d_input(xs[0].shape, w.float(), dys[0].float(), st, pad)
    for on in (0, 1):
        ext.set_cluster_multicast(on)
        y = ops.conv_fprop(xs[0], w, st, pad)
        dx = ops.conv_dgrad(dys[0], w, xs[0].shape, st, pad)
        torch.cuda.synchronize()
        r[f"fprop_err_mc{on}"] = round(float((y.float() - yref).abs().max() / yref.abs().max()), 4)
        r[f"dgrad_err_mc{on}"] = round(float((dx.float() - dref).abs().max() / dref.abs().max()), 4)
        r[f"fprop_us_mc{on}"] = timeit(lambda i: (lambda: ops.conv_fprop(xs[i], w, st, pad)))
        r[f"dgrad_us_mc{on}"] = timeit(lambda i: (lambda: ops.conv_dgrad(dys[i], w, xs[i].shape, st, pad)))
    rows.append(r)
    print(json.dumps(r), flush=True)
json.dump(rows, open("gpurun_out/mc_check.json", "w"), indent=1)
