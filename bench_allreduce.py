#!/usr/bin/env python
"""Gradient all-reduce + optimiser-update sweep (BASELINE.json config #5).

For bucket sizes 64 KB .. 1 GB (fp32 gradient bytes) compares, at N GPUs of one node
(launch under ``torch.distributed.run``, one rank per GPU):

  fused/p2p      our kernel, two-shot, peer loads/stores          (csrc/fused_update.cu)
  fused/nvls     our kernel, two-shot, multimem.ld_reduce / multimem.st
  fused/oneshot  our kernel, one-shot (small buckets)
  nccl+opt       baseline: NCCL all_reduce(SUM) + scale + a separate fused torch Momentum update kernel

Device-timed with CUDA events, max over ranks, 5 warm-up + 20 timed launches per point.
Bytes that must cross NVLink per rank for a two-shot all-reduce of G gradient bytes followed by a
broadcast of the updated bf16 weights:  in: G*(N-1)/N, out: (G/2)*(N-1)/N  (fp32 grads in, bf16 weights out).
Roofline time = max(in, out) / 770 GB/s (measured peer-copy bandwidth per direction, B200_PROFILING.md); the
fraction against the nominal 900 GB/s per direction (BASELINE.json) is reported next to it (``frac_of_900``);
``bus_gbs`` follows the NCCL convention 2*(N-1)/N * G / t so it is comparable with nccl-tests.
Prints one JSON line per (variant, size) and a final summary line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-kb", type=int, default=64)
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    import fluxdist as fd
    from fluxdist import ops
    from fluxdist.parallel.native_comm import MAX_GRID, MODE_ONE_SHOT, MODE_TWO_SHOT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    ext = ops.extension()
    gname = dist.group.WORLD.group_name
    LINK = 770e9

    sizes = []
    b = args.min_kb * 1024
    while b <= args.max_mb * 1024 * 1024:
        sizes.append(b)
        b *= 4
    nmax = sizes[-1] // 4
    align = 2048 * world
    nmax = (nmax + align - 1) // align * align

    g = symm.empty(nmax, dtype=torch.float32, device=dev)
    w16 = symm.empty(nmax, dtype=torch.bfloat16, device=dev)
    flags = symm.empty(MAX_GRID * 8, dtype=torch.int32, device=dev)
    g.normal_()
    w16.zero_()
    flags.zero_()
    hg, hw, hf = symm.rendezvous(g, gname), symm.rendezvous(w16, gname), symm.rendezvous(flags, gname)
    w32 = torch.zeros(nmax, device=dev)
    s0 = torch.zeros(nmax, device=dev)
    epoch = torch.zeros(MAX_GRID, dtype=torch.int32, device=dev)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    done = torch.zeros(1, dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    has_mc = bool(getattr(hg, "multicast_ptr", 0))
    torch.cuda.synchronize()
    dist.barrier()

    def plan(n, mode, nvls, unroll=2):
        p = ext.FusedPlan(world, rank, 1, mode, 1 if nvls else 0, n, n, list(hg.buffer_ptrs), w32.data_ptr(),
                          s0.data_ptr(), 0, [], list(hw.buffer_ptrs), list(hf.buffer_ptrs), epoch.data_ptr(),
                          step.data_ptr(), done.data_ptr(), status.data_ptr(), int(hg.multicast_ptr) if nvls else 0, 0,
                          int(hw.multicast_ptr) if nvls else 0, MAX_GRID, local)
        p.set_hyper([0.01, 0.9, 0, 0, 0], 1.0)
        p.set_unroll(unroll)
        return p

    def timed(fn, iters):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b_.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b_) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3  # s

    gn = torch.empty(nmax, device=dev)
    rows = []
    for nbytes in sizes:
        n = nbytes // 4
        n = max(align, (n + align - 1) // align * align)
        G = n * 4
        variants = [("fused/p2p", lambda p=plan(n, MODE_TWO_SHOT, False): p.launch())]
        variants.append(("fused/p2p/u4", lambda p=plan(n, MODE_TWO_SHOT, False, 4): p.launch()))
        if has_mc:
            variants.append(("fused/nvls", lambda p=plan(n, MODE_TWO_SHOT, True): p.launch()))
            variants.append(("fused/nvls/u4", lambda p=plan(n, MODE_TWO_SHOT, True, 4): p.launch()))
        if G <= 8 * 1024 * 1024:
            variants.append(("fused/oneshot", lambda p=plan(n, MODE_ONE_SHOT, False): p.launch()))

        def nccl_fn(n=n):
            gv = gn[:n]
            gv.copy_(g[:n])
            dist.all_reduce(gv)
            gv.mul_(1.0 / world)
            s0[:n].mul_(0.9).add_(gv, alpha=-0.01)
            w32[:n].add_(s0[:n])
            w16[:n].copy_(w32[:n])

        variants.append(("nccl+opt", nccl_fn))
        for name, fn in variants:
            t = timed(fn, args.iters)
            link_bytes = G * (world - 1) / world
            roof = link_bytes / LINK if world > 1 else G * 4.5 / 6.5e12
            rec = {"variant": name, "bytes": G, "n_gpus": world, "us": t * 1e6,
                   "bus_gbs": (2 * (world - 1) / world * G / t / 1e9) if world > 1 else None,
                   "algo_gbs": G / t / 1e9, "roofline_us": roof * 1e6, "frac_of_roofline": roof / t,
                   "frac_of_900": (link_bytes / 900e9 / t) if world > 1 else None}
            rows.append(rec)
            if rank == 0:
                print(json.dumps(rec), flush=True)
    st = status.tolist()
    if rank == 0:
        best = {}
        for r in rows:
            if r["variant"].startswith("fused"):
                k = r["bytes"]
                if k not in best or r["us"] < best[k]["us"]:
                    best[k] = r
        print(json.dumps({"summary": "fused all-reduce+update sweep", "n_gpus": world, "nvls": has_mc,
                          "barrier_timeout": bool(st[1]),
                          "best_fraction_of_roofline": max(r["frac_of_roofline"] for r in best.values()),
                          "largest": best[max(best)]}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
