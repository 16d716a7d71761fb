"""Debug mode of the fused all-reduce + update: poison the buckets, run the exchange, verify data AND protocol state
("a debug mode for the fused all-reduce that poisons buckets and verifies flag/epoch protocol").

The reference has nothing of the kind — its gradient exchange (``src/ddp_tasks.jl:59-109,163-172``) is guarded by two
full barriers and its consistency check ``ensure_synced`` (``:115-126``) is never called.  Here the exchange is ONE kernel
per bucket with device-side barriers (``csrc/fused_update.cu``), so a protocol slip (a CTA that passes a barrier one
epoch early, a slice nobody reduces, a broadcast that misses a replica) would not crash: it would silently train on
partly stale values.  ``verify_exchange`` makes such a slip loud:

1. every replica's state is snapshotted (weights, moments, BN buffers, step counters);
2. for ``rounds`` rounds the gradient arena of rank ``r`` is filled with ``(r + 1) * k * pattern`` (a deterministic,
   non-zero, index-dependent pattern, different every round so that a value left over from the previous round cannot
   pass), master weights and moments are zeroed, and every OUTPUT of the kernel is poisoned with NaN first: the bf16
   compute weights everywhere and, on the sharded (two-shot) native path, the fp32 weights outside the rank's own
   slice;
3. the real ``reduce_and_update`` runs on all replicas (threads in task mode, collectively in process mode);
4. checks: every parameter element of the compute weights equals the optimiser rule applied to the known mean
   ``k * pattern * (N + 1) / 2`` (so every slice was reduced over ALL ranks, scaled by 1/N, updated once and broadcast
   to this replica; a surviving NaN = an element nobody wrote); native path: status words clean (no barrier timeout,
   no non-finite flag), the device step counter of every bucket advanced by exactly one per round, and the epoch words
   of all CTAs of a bucket agree (each CTA went through the same number of barrier phases);
5. the snapshot is restored, so the call is side-effect free.

Works with every communicator (CPU threads / gloo / native), which is what the CPU tests exercise; on GPUs it runs the
production kernel, plans and buffers.  ``FLUXDIST_DEBUG_EXCHANGE=1`` makes ``prepare_training`` and the process-mode
workers run it once before training.
"""
from __future__ import annotations

import threading
from typing import Dict, List

import torch

from ..devices import device_ctx
from ..optim.rules import flat_step

__all__ = ["verify_exchange", "ExchangeError"]


class ExchangeError(RuntimeError):
    pass


def _pattern(n: int, device) -> torch.Tensor:
    """Deterministic, never-zero, index-dependent fp32 pattern with magnitudes in [0.25, 1.25) and alternating sign."""
    i = torch.arange(n, dtype=torch.int64, device=device)
    mag = 0.25 + ((i * 2654435761) % 1021).to(torch.float32) / 1021.0
    return torch.where((i & 1) == 0, mag, -mag)


def _param_mask(arena) -> torch.Tensor:
    m = torch.zeros(arena.g32.numel(), dtype=torch.bool, device=arena.g32.device)
    for off, p in zip(arena.offsets, arena.params):
        m[off: off + p.numel()] = True
    return m


def _run_all(replicas) -> None:
    """reduce_and_update on every local replica; concurrently when there are several (the CPU communicators block in
    host barriers, co-located native replicas rendezvous on a host barrier too)."""
    if len(replicas) == 1:
        r = replicas[0]
        with device_ctx(r.dev):
            r.reduce_and_update()
        return
    errs: List[BaseException] = []

    def work(r):
        try:
            with device_ctx(r.dev):
                r.reduce_and_update()
        except BaseException as e:  # noqa: BLE001 - reported to the caller below
            errs.append(e)
            comm = r.group.comm
            if hasattr(comm, "abort"):
                comm.abort()

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in replicas]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]


def verify_exchange(group, rounds: int = 2, raise_on_error: bool = True) -> Dict[str, object]:
    """Poisoned-bucket self test of ``group``'s gradient exchange (see module docstring).  Collective: in process mode
    every rank must call it.  Returns a report dict; raises :class:`ExchangeError` on the first violated check unless
    ``raise_on_error`` is false (then ``report["errors"]`` lists them)."""
    reps = group.replicas
    comm = group.comm
    native = getattr(comm, "name", "") == "native"
    errors: List[str] = []
    report: Dict[str, object] = {"comm": getattr(comm, "name", type(comm).__name__), "rounds": rounds,
                                 "replicas": len(reps), "checked_elems": 0, "max_rel_err": 0.0, "errors": errors}

    def fail(msg: str) -> None:
        errors.append(msg)
        if raise_on_error:
            raise ExchangeError(msg)

    snaps = [r.snapshot_state() for r in reps]
    launched = [set(r._launched) for r in reps]
    epochs0 = {}
    try:
        masks = [_param_mask(r.arena) for r in reps]
        pats = [_pattern(r.arena.g32.numel(), r.arena.g32.device) for r in reps]
        for k in range(1, rounds + 1):
            for r, pat in zip(reps, pats):
                a = r.arena
                with device_ctx(r.dev), torch.no_grad():
                    a.g32.copy_(pat * float((r.rank + 1) * k))
                    a.w32.zero_()
                    for s in a.slots:
                        s.zero_()
                    if a.w16 is not None:
                        a.w16.fill_(float("nan"))
                    if native and r.world > 1 and comm._two_shot(r):
                        own = torch.zeros(a.w32.numel(), dtype=torch.bool, device=a.w32.device)
                        for lo, hi, rk in a.owner_ranges(r.world):
                            if rk == r.rank:
                                own[lo:hi] = True
                        a.w32[~own] = float("nan")     # outputs of the fp32 broadcast + slices this rank never reads
                    r.step_count = 0
                    r._launched = set()
                    if hasattr(comm, "set_step"):
                        comm.set_step(r, 0)
                    if native:
                        comm.status(r)  # read-and-clear: flags raised before this round (kept in the snapshot) are not ours
                    if native and k == 1:
                        aux = comm.plans[id(r)][1]
                        epochs0[id(r)] = aux["epoch"].clone()
            _run_all(reps)
            for r, pat, mask in zip(reps, pats, masks):
                a = r.arena
                with device_ctx(r.dev), torch.no_grad():
                    if r.dev.is_cuda:
                        torch.cuda.synchronize(r.dev.torch_device)
                    mean = pat * (float(k) * (r.world + 1) / 2.0)
                    exp = torch.zeros_like(mean)
                    flat_step(r.opt, exp, mean, [torch.zeros_like(mean) for _ in a.slots], 1)
                    got = (a.w16 if a.w16 is not None else a.w32).float()
                    lowp = a.w16 is not None
                    g, e = got[mask], exp[mask]
                    bad = ~torch.isfinite(g)
                    if bool(bad.any()):
                        fail(f"round {k}, rank {r.rank}: {int(bad.sum())} parameter elements of the compute weights were "
                             f"never written (still NaN), first at arena index {int(mask.nonzero()[bad.nonzero()[0, 0], 0])}")
                        continue
                    rel = (g - e).abs() / (e.abs() + 1e-6)
                    tol = 1.2e-2 if lowp else 2e-4   # bf16 rounding of the broadcast weights / fp32 summation order
                    worst = float(rel.max()) if rel.numel() else 0.0
                    report["max_rel_err"] = max(float(report["max_rel_err"]), worst)
                    report["checked_elems"] = int(report["checked_elems"]) + int(g.numel())
                    if worst > tol:
                        j = int(rel.argmax())
                        fail(f"round {k}, rank {r.rank}: compute weight differs from rule(mean over {r.world} ranks): "
                             f"got {float(g[j]):.6g}, expected {float(e[j]):.6g} (rel {worst:.3g}) at masked index {j} — "
                             "a slice was not reduced over all ranks, scaled, or broadcast")
                    if native and r.world > 1 and comm._two_shot(r):
                        # fp32 outputs: the rank's own master slice and the fp32-broadcast tail must hold the result too
                        bnd = int(getattr(r, "comm_info", {}).get("fp32_bcast_from", a.w32.numel()))
                        chk = torch.zeros_like(mask)
                        for lo, hi, rk in a.owner_ranges(r.world):
                            if rk == r.rank:
                                chk[lo:hi] = True
                        chk[bnd:] = True
                        chk &= mask
                        g32, e32 = a.w32[chk], exp[chk]
                        if not bool(torch.isfinite(g32).all()):
                            fail(f"round {k}, rank {r.rank}: fp32 master / fp32-broadcast region has unwritten elements")
                        elif g32.numel() and float(((g32 - e32).abs() / (e32.abs() + 1e-6)).max()) > 2e-4:
                            fail(f"round {k}, rank {r.rank}: fp32 master / fp32-broadcast region differs from the expected update")
                    if native:
                        nan, tmo = comm.status(r)
                        if tmo:
                            fail(f"round {k}, rank {r.rank}: a device-side barrier timed out")
                        if nan:
                            fail(f"round {k}, rank {r.rank}: non-finite flag raised on finite gradients")
                        plans, aux, _ = comm.plans[id(r)]
                        steps = aux["step"][: len(plans)].tolist()
                        if any(s != 1 for s in steps):
                            fail(f"round {k}, rank {r.rank}: device step counters {steps} (expected 1 per bucket)")
        if native:
            from .native_comm import MAX_GRID

            deltas = []
            for r in reps:
                plans, aux, _ = comm.plans[id(r)]
                d = (aux["epoch"] - epochs0[id(r)]).tolist()
                for b, plan in enumerate(plans):
                    seg = d[b * MAX_GRID: b * MAX_GRID + int(plan.grid)]
                    if len(set(seg)) != 1:
                        fail(f"rank {r.rank}, bucket {b}: CTAs disagree on the barrier epoch: {sorted(set(seg))}")
                    if r.world > 1 and seg and seg[0] <= 0:
                        fail(f"rank {r.rank}, bucket {b}: barrier epoch did not advance")
                    deltas.append((b, seg[0] if seg else 0))
            per_bucket: Dict[int, set] = {}
            for b, v in deltas:
                per_bucket.setdefault(b, set()).add(v)
            for b, vs in per_bucket.items():
                if len(vs) != 1:
                    fail(f"bucket {b}: replicas disagree on the number of barrier phases: {sorted(vs)}")
            report["epoch_advance_per_bucket"] = {b: sorted(vs)[0] for b, vs in per_bucket.items() if vs}
    finally:
        for r, s, l in zip(reps, snaps, launched):
            r.restore_state(s)
            r._launched = l
            with device_ctx(r.dev):
                r.arena.g32.zero_()
        if hasattr(comm, "reset"):
            comm.reset()
    report["ok"] = not errors
    return report
