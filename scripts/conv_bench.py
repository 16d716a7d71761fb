"""Per-layer timing of the ResNet-34 3x3/s1 conv shapes (batch 96): cuDNN vs im2col-TMA vs halo-tile kernels.
Each candidate is captured into a CUDA graph of 12 launches over 4 rotating buffer sets (working set > L2, so every
launch sees a cold L2 for its activations) and timed with CUDA events over 5 replays: host launch overhead is excluded."""
import json, sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
import torch.nn.functional as F
BF = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
shapes = [(56, 64, 64), (28, 128, 128), (14, 256, 256), (7, 512, 512)]
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
for hw, C, K in shapes:
    xs = [cl(torch.randn(B, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    dys = [cl(torch.randn(B, K, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    w = cl((torch.randn(K, C, 3, 3, device="cuda") * 0.05).to(BF))
    gb = torch.zeros(K, C, 3, 3, device="cuda").contiguous(memory_format=torch.channels_last)
    flops = 2.0 * B * hw * hw * C * K * 9
    r = {"hw": hw, "C": C, "K": K, "gflop": round(flops / 1e9, 1)}
    def rec(name, mk):
        try:
            r[name] = round(timeit(mk), 1)
        except Exception as e:
            r[name] = None
    rec("cudnn_fprop", lambda i: (lambda: F.conv2d(xs[i], w, None, 1, 1)))
    rec("cudnn_dgrad", lambda i: (lambda: torch.nn.grad.conv2d_input(xs[i].shape, w, dys[i], 1, 1)))
    rec("cudnn_wgrad", lambda i: (lambda: torch.nn.grad.conv2d_weight(xs[i], w.shape, dys[i], 1, 1)))
    rec("im2col_fprop", lambda i: (lambda: ops.conv_fprop(xs[i], w, 1, 1)))
    rec("im2col_dgrad", lambda i: (lambda: ops.conv_dgrad(dys[i], w, xs[i].shape, 1, 1)))
    rec("im2col_wgrad", lambda i: (lambda: ops.conv_wgrad(dys[i], xs[i], w.shape, 1, 1, out=gb)))
    for v in (1, 2, 3):
        rec(f"halo_fprop_v{v}", lambda i: (lambda: ops.conv3x3_fprop(xs[i], w, stats=False, variant=v)))
        rec(f"halo_fprop_stats_v{v}", lambda i: (lambda: ops.conv3x3_fprop(xs[i], w, stats=True, variant=v)))
        rec(f"halo_dgrad_v{v}", lambda i: (lambda: ops.conv3x3_dgrad(dys[i], w, variant=v)))
    rec("halo_wgrad", lambda i: (lambda: ops.conv3x3_wgrad(dys[i], xs[i], out=gb)))
    rows.append(r)
    print(json.dumps(r), flush=True)
json.dump(rows, open("gpurun_out/conv_bench.json", "w"), indent=1)
