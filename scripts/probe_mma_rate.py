import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
e = ops.extension()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for bn in (64, 128, 256):
    for shift, sbo in [(0, 1024), (1, 1024), (0, 1280), (11, 1280), (0, 2048)]:
        e.probe_mma_rate(out, bn, 64, shift, sbo); torch.cuda.synchronize()
        e.probe_mma_rate(out, bn, 256, shift, sbo); torch.cuda.synchronize()
        cyc = out.item() / (256 * 4)
        print(f"N={bn:3d} shift={shift:2d} sbo={sbo}: {cyc:.1f} cycles per tcgen05.mma (128xNx16)  ideal {128*bn/256:.0f}")
