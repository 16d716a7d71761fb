import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
e = ops.extension()
out = torch.zeros(148, dtype=torch.int64, device="cuda")
reps = 9 * 32
for bn in (64, 128):
    for grid in (1, 148):
        for flavor, name in [(0, "same acc, no commits"), (1, "alternate acc"), (3, "alt acc + commit/tile"), (7, "alt acc + commit + tap shifts + per-tap B"), (11, "alt acc + commit + WAIT per tile")]:
            e.probe_mma_rate(out, bn, reps, 0, 1280, grid, flavor); torch.cuda.synchronize()
            e.probe_mma_rate(out, bn, reps, 0, 1280, grid, flavor); torch.cuda.synchronize()
            cyc = out[:grid].float().mean().item() / (reps * 4)
            print(f"N={bn:3d} grid={grid:3d} {name:45s}: {cyc:6.1f} cycles per MMA", flush=True)
