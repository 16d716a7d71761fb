"""Sustained tcgen05.mma rate per operand form and N (csrc/probes/probe_mma_rate2.cu), 1 SM and all SMs."""
import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
e = ops.probe_extension()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
reps = 512
names = {0: "A,B K-major (fprop form)", 1: "B MN-major (dgrad form)", 2: "A,B MN-major (wgrad form)"}
for cta2 in (False, True):
    for bn in (64, 128, 256):
        for form in (0, 1, 2):
            for grid in (2, 148):
                e.probe_mma_rate2(out, bn, form, cta2, reps, grid); torch.cuda.synchronize()
                e.probe_mma_rate2(out, bn, form, cta2, reps, grid); torch.cuda.synchronize()
                v = out[:grid].float()
                cyc = v[v > 0].mean().item() / (reps * 4)
                m = 256 if cta2 else 128
                floor = (m * bn) / (256 * (2 if cta2 else 1))
                print(f"{'cta_group::2 M=256' if cta2 else 'cta_group::1 M=128'} N={bn:3d} grid={grid:3d} {names[form]:28s}: "
                      f"{cyc:6.1f} clk/MMA (floor {floor:.0f}) -> {100 * floor / cyc:5.1f}% of peak", flush=True)
