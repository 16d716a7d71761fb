"""L2 -> shared-memory fill bandwidth of tiled TMA loads (csrc/probes/probe_tma_bw.cu).
Prints bytes/clk chip-wide for: every CTA its own L2-resident region, pairs / all CTAs on the same region, and
2-CTA-cluster multicast.  clock64 counts SM cycles, so B/clk is independent of the clock the GPU settles at."""
import sys, torch
sys.path.insert(0, ".")
import fluxdist
from fluxdist import ops
e = ops.probe_extension()
GRID = 148
BOXES = 24            # 24 boxes x 16 KB = 384 KB per region; 148 regions = 57 MB (L2-resident after the first pass)
rows = GRID * BOXES * 128
buf = torch.randn(rows, 64, device="cuda").to(torch.bfloat16)
cyc = torch.zeros(GRID, dtype=torch.int64, device="cuda")
for share, mc, name in [(1, False, "distinct region per CTA"), (2, False, "2 CTAs per region (unicast)"),
                        (4, False, "4 CTAs per region (unicast)"), (148, False, "all CTAs same region (unicast)"),
                        (1, True, "2-CTA cluster, half box each, multicast"), (74, True, "all clusters same region, multicast")]:
    for reps in (2, 16):
        e.probe_tma_bw(buf, cyc, GRID, BOXES, share, reps, mc); torch.cuda.synchronize()
    c = cyc.float().max().item()
    per_cta = BOXES * 16 * 16384 / c
    print(f"{name:45s}: {per_cta:7.1f} B/clk per SM delivered, {per_cta * GRID:9.0f} B/clk chip-wide", flush=True)
# fewer CTAs: a per-SM port limit keeps B/clk per SM constant, a chip-wide (L2) limit lets it rise
for g in (8, 32, 74):
    for reps in (2, 16):
        e.probe_tma_bw(buf, cyc, g, BOXES, 1, reps, False); torch.cuda.synchronize()
    c = cyc[:g].float().max().item()
    per_cta = BOXES * 16 * 16384 / c
    print(f"grid={g:3d} distinct regions                        : {per_cta:7.1f} B/clk per SM delivered, {per_cta * g:9.0f} B/clk chip-wide", flush=True)
