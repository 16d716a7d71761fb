import sys, torch
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
for hw, C, K, v in [(56, 64, 64, 1), (28, 128, 128, 3)]:
    xs = [cl(torch.randn(96, C, hw, hw, device="cuda").to(BF)) for _ in range(NSET)]
    w = cl((torch.randn(K, C, 3, 3, device="cuda") * 0.05).to(BF))
    for dbg in (0, 1, 4, 8, 9, 12, 13, 15):
        t = timeit(lambda i: (lambda: ops.conv3x3_fprop(xs[i], w, stats=False, variant=v + (dbg << 9))))
        print(f"hw={hw} C={C} v{v} dbg={dbg:04b} (1=nostore 2=noload 4=nomma 8=noldtm): {t:.1f} us", flush=True)
