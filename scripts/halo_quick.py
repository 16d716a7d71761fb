import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
BF = torch.bfloat16
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n * 1e3
for hw, C, K in [(56, 64, 64), (28, 128, 128), (14, 256, 256)]:
    x = cl(torch.randn(96, C, hw, hw, device="cuda").to(BF)); w = cl((torch.randn(K, C, 3, 3, device="cuda") * 0.05).to(BF))
    out = {}
    for v in (1, 2, 3):
        for st in (False, True):
            try:
                out[f"fprop_v{v}_stats{int(st)}"] = round(timeit(lambda: ops.conv3x3_fprop(x, w, stats=st, variant=v)), 1)
            except Exception as e:
                pass
        try:
            out[f"dgrad_v{v}"] = round(timeit(lambda: ops.conv3x3_dgrad(x, w, variant=v)), 1)
        except Exception:
            pass
    print(hw, C, K, out, flush=True)
