"""Round-2 per-layer conv benchmark: cuDNN vs this repo's kernels on the ResNet-34 3x3/s1 shapes (batch 96) plus the
strided convs.  Same method as conv_bench.py (CUDA graph of 12 launches over 4 rotating buffer sets = cold L2 for the
activations, CUDA events over 5 replays).  Writes gpurun_out/conv_bench_r2.json."""
import json, os, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
import torch.nn.functional as F
BF = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
ext = ops.extension()
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
rows = []
shapes = [(56, 64, 64, 3, 1), (28, 128, 128, 3, 1), (14, 256, 256, 3, 1), (7, 512, 512, 3, 1),
          (56, 64, 128, 3, 2), (28, 128, 256, 3, 2), (14, 256, 512, 3, 2), (56, 64, 128, 1, 2)]
for hw, C, K, R, st in shapes:
    pad = R // 2
    ho = (hw + 2 * pad - R) // st + 1
    xs = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    dys = [cl(torch.randn(B, K, ho, ho, device="cuda").to(BF)) for _ in range(NSET)]
    w = cl((torch.randn(K, C, R, R, device="cuda") * 0.05).to(BF))
    gb = torch.zeros(K, C, R, R, device="cuda").contiguous(memory_format=torch.channels_last)
    sm, sq = torch.zeros(K, device="cuda"), torch.zeros(K, device="cuda")
    flops = 2.0 * B * ho * ho * C * K * R * R
    r = {"hw": hw, "C": C, "K": K, "R": R, "stride": st, "gflop": round(flops / 1e9, 1)}
    def rec(name, mk):
        try:
            r[name] = round(timeit(mk), 1)
        except Exception as e:
            r[name] = None
            r[name + "_err"] = str(e)[:80]
    rec("cudnn_fprop", lambda i: (lambda: F.conv2d(xs[i], w, None, st, pad)))
    rec("cudnn_dgrad", lambda i: (lambda: torch.nn.grad.conv2d_input(xs[i].shape, w, dys[i], st, pad)))
    rec("cudnn_wgrad", lambda i: (lambda: torch.nn.grad.conv2d_weight(xs[i], w.shape, dys[i], st, pad)))
    for bn, mt in [(0, 0), (64, 1), (64, 2), (128, 1), (128, 2), (256, 1)]:
        tag = "auto" if bn == 0 else f"{bn}x{mt}"
        if bn and (K % bn or C % bn):
            continue
        ext.set_tile_shape(bn, mt)
        rec(f"fprop_{tag}", lambda i: (lambda: ops.conv_fprop(xs[i], w, st, pad)))
        rec(f"fprop_stats_{tag}", lambda i: (lambda: ops.conv_fprop(xs[i], w, st, pad, stats=(sm, sq))))
        rec(f"dgrad_{tag}", lambda i: (lambda: ops.conv_dgrad(dys[i], w, xs[i].shape, st, pad)))
    ext.set_tile_shape(0, 0)
    for sp in (0, 1, 2, 4, 8, 16, 32):
        rec(f"wgrad_im2col_s{sp}", lambda i: (lambda: ops.conv_wgrad(dys[i], xs[i], w.shape, st, pad, out=gb, splits=sp)))
    if R == 3 and st == 1:
        for v in (1, 3):
            rec(f"halo_fprop_v{v}", lambda i: (lambda: ops.conv3x3_fprop(xs[i], w, stats=False, variant=v)))
            # pre-zeroed accumulators, as in the model (bn_apply re-zeroes them): no memset nodes in the timed graph
            rec(f"halo_fprop_stats_v{v}", lambda i: (lambda: ops.conv3x3_fprop(xs[i], w, stats=(sm, sq), variant=v)))
            rec(f"halo_dgrad_v{v}", lambda i: (lambda: ops.conv3x3_dgrad(dys[i], w, variant=v)))
        rec("halo_wgrad", lambda i: (lambda: ops.conv3x3_wgrad(dys[i], xs[i], out=gb)))
    rows.append(r)
    print(json.dumps(r), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/conv_bench_r2.json", "w"), indent=1)
