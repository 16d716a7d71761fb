"""NativeComm: binds FusedReplicas to the hand-written sm_100a all-reduce+update kernel
(``csrc/fused_update.cu``) over peer-mapped memory.

Two rendezvous paths give every replica the addresses of every peer's arenas:

* **thread mode** (task-based parallelism, one process, N GPUs): plain device
  allocations + ``cudaDeviceEnablePeerAccess``; with UVA a peer's ``data_ptr()`` is
  directly dereferenceable from a kernel on any other GPU of the process.
* **process mode** (one process per GPU, ``torch.distributed`` NCCL group): the
  gradient / weight-out / signal buffers are ``torch.distributed._symmetric_memory``
  allocations; ``rendezvous`` yields ``buffer_ptrs`` (P2P mappings) and, when the
  NVSwitch supports it, ``multicast_ptr`` for the NVLS variant
  (``multimem.ld_reduce`` / ``multimem.st``).

Variant choice per bucket (static, so the optimiser-state layout is fixed —
SURVEY §7.4-4): ``one_shot`` below ``one_shot_max_bytes`` (latency-bound: one
barrier pair, redundant math, no broadcast), ``two_shot`` above; NVLS when a
multicast mapping exists and ``FLUXDIST_NVLS != 0``.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
from torch import nn

from ..devices import as_device, device_ctx
from ..optim import Rule
from ..optim.rules import RULE_ADAM, RULE_ADAMW, RULE_DESCENT, RULE_MOMENTUM, RULE_NESTEROV
from .fused import FlatArena, FusedReplica

MODE_TWO_SHOT, MODE_ONE_SHOT = 0, 1
MAX_GRID = 296


def _hyper(opt: Rule) -> List[float]:
    if opt.rule_id == RULE_DESCENT:
        return [opt.eta, 0, 0, 0, 0]
    if opt.rule_id in (RULE_MOMENTUM, RULE_NESTEROV):
        return [opt.eta, opt.rho, 0, 0, 0]
    return [opt.eta, opt.beta[0], opt.beta[1], opt.eps, getattr(opt, "decay", 0.0)]


def _lowp_only_boundary(model: nn.Module, arena: FlatArena) -> int:
    """Offset from which parameters are consumed in fp32 by the forward pass (BN affine,
    biases).  Below it every parameter is a conv/FC weight read through its bf16 copy, so the
    kernel broadcasts bf16 only.  Returns 0 (broadcast everything in fp32 too) for models with
    layers that read fp32 ``p.data`` directly."""
    from ..nn.layers import ConvBN, Head, get_backend

    if arena.w16 is None or get_backend() == "torch":
        return 0
    lowp_ids = set()
    for m in model.modules():
        if isinstance(m, ConvBN):  # every ConvBN reads its filter through the bf16 copy (tcgen05 or stem path)
            lowp_ids.add(id(m.conv.weight))
        elif isinstance(m, Head):
            lowp_ids.add(id(m.fc.weight))
    boundary = 0
    for p, off in zip(arena.params, arena.offsets):
        if id(p) in lowp_ids and off == boundary:
            boundary = off + ((p.numel() + 7) // 8) * 8
        else:
            break
    return boundary


class NativeComm:
    name = "native"

    def __init__(self, world: int, mode: Optional[int] = None, one_shot_max_bytes: int = 256 * 1024):
        self.world = world
        self.mode_override = mode
        self.one_shot_max_bytes = one_shot_max_bytes
        self.plans = {}
        self._keep = []  # keeps symmetric-memory handles and helper tensors alive

    # ------------------------------------------------------------------ construction: thread mode
    @classmethod
    def build_thread_replicas(cls, models: Sequence[nn.Module], devs, opt: Rule, lowp, mode: Optional[int] = None,
                              **kw) -> List[FusedReplica]:
        from .. import ops

        ext = ops.extension()
        n = len(models)
        comm = cls(n, mode, **kw)
        cu = [d.cuda_index for d in devs]
        for a in set(cu):
            for b in set(cu):
                if a != b:
                    ext.enable_peer_access(a, b)
        reps = []
        for i, (m, d) in enumerate(zip(models, devs)):
            rep = FusedReplica(m, d, opt, n, i, lowp, sort_lowp_first=True)
            rep._native_comm = comm
            reps.append(rep)
        share = max(cu.count(c) for c in set(cu))
        if share > 1:
            # virtual replicas on one physical GPU (tests, 1-GPU development): their kernels spin on each other, so
            # (a) all of them must be co-resident -> bound the grid; (b) nothing a replica still has to do before its
            # launch may need the context to go idle (lazy module loading does) -> host rendezvous before the launch
            import threading

            comm._shared_gpu = threading.Barrier(n)
            comm.max_grid = max(1, 148 // share)
        aux = []
        for rep in reps:
            with device_ctx(rep.dev):
                tdev = rep.dev.torch_device
                aux.append(dict(flags=torch.zeros(MAX_GRID * 8, dtype=torch.int32, device=tdev),
                                epoch=torch.zeros(MAX_GRID, dtype=torch.int32, device=tdev),
                                step=torch.zeros(1, dtype=torch.int32, device=tdev),
                                done=torch.zeros(1, dtype=torch.int32, device=tdev),
                                status=torch.zeros(2, dtype=torch.int32, device=tdev)))
        for i, rep in enumerate(reps):
            a = rep.arena
            comm._make_plan(rep, g_ptrs=[r.arena.g32.data_ptr() for r in reps],
                            w32_ptrs=[r.arena.w32.data_ptr() for r in reps],
                            w16_ptrs=[(r.arena.w16.data_ptr() if r.arena.w16 is not None else 0) for r in reps],
                            flag_ptrs=[x["flags"].data_ptr() for x in aux], aux=aux[i], mc=(0, 0, 0))
        comm._keep.append(aux)
        comm.replicas = reps
        return reps

    # ------------------------------------------------------------------ construction: process mode
    @classmethod
    def build_process_replica(cls, model: nn.Module, dev, opt: Rule, lowp, world: int, rank: int, group=None,
                              mode: Optional[int] = None, **kw) -> FusedReplica:
        import torch.distributed as dist

        comm = cls(world, mode, **kw)
        tdev = dev.torch_device
        mc = (0, 0, 0)
        if world == 1:
            rep = FusedReplica(model, dev, opt, 1, 0, lowp, sort_lowp_first=True)
            g_ptrs, w32_ptrs = [rep.arena.g32.data_ptr()], [rep.arena.w32.data_ptr()]
            w16_ptrs = [rep.arena.w16.data_ptr() if rep.arena.w16 is not None else 0]
            flags = torch.zeros(MAX_GRID * 8, dtype=torch.int32, device=tdev)
            flag_ptrs = [flags.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm

            gname = (group or dist.group.WORLD).group_name

            def alloc(nel, dt):
                t = symm.empty(nel, dtype=dt, device=tdev)
                t.zero_()
                return t

            rep = FusedReplica(model, dev, opt, world, rank, lowp, alloc=alloc, sort_lowp_first=True)
            a = rep.arena
            dist.broadcast(a.w32, src=0, group=group)  # replicas start identical
            if a.w16 is not None:
                a.w16.copy_(a.w32)
            flags = symm.empty(MAX_GRID * 8, dtype=torch.int32, device=tdev)
            flags.zero_()
            hg = symm.rendezvous(a.g32, gname)
            hw = symm.rendezvous(a.w32, gname)
            hf = symm.rendezvous(flags, gname)
            g_ptrs, w32_ptrs, flag_ptrs = list(hg.buffer_ptrs), list(hw.buffer_ptrs), list(hf.buffer_ptrs)
            w16_ptrs = [0] * world
            h16 = None
            if a.w16 is not None:
                h16 = symm.rendezvous(a.w16, gname)
                w16_ptrs = list(h16.buffer_ptrs)
            # measured: on 8 GPUs the in-switch reduction wins at every size (bench_allreduce.py: 141 vs 215 us at
            # 64 MB); on 2 GPUs a direct peer read of half the bucket beats the round trip through the switch from
            # 4 MB up (109 vs 157 us at 64 MB), and on 4 GPUs the ResNet-34 step is 5.46 ms with peer loads vs 5.78 ms
            # with NVLS (profiles/bench_native_x4_nvls{0,1}.json).  FLUXDIST_NVLS=0/1 forces it.
            env = os.environ.get("FLUXDIST_NVLS", "auto")
            want_nvls = (env != "0") if env in ("0", "1") else (world >= 8 or a.padded * 4 < (4 << 20))
            use_nvls = want_nvls and bool(getattr(hg, "multicast_ptr", 0))
            if use_nvls:
                mc = (int(hg.multicast_ptr), int(hw.multicast_ptr), int(h16.multicast_ptr) if h16 is not None else 0)
            comm._keep += [hg, hw, hf, h16, flags]
            torch.cuda.synchronize(tdev)
            dist.barrier(group=group)
        rep._native_comm = comm
        aux = dict(flags=flags, epoch=torch.zeros(MAX_GRID, dtype=torch.int32, device=tdev),
                   step=torch.zeros(1, dtype=torch.int32, device=tdev),
                   done=torch.zeros(1, dtype=torch.int32, device=tdev),
                   status=torch.zeros(2, dtype=torch.int32, device=tdev))
        comm._make_plan(rep, g_ptrs, w32_ptrs, w16_ptrs, flag_ptrs, aux, mc)
        comm._keep.append(aux)
        comm.replicas = [rep]
        comm.group = group
        return rep

    # ------------------------------------------------------------------ plan
    def _make_plan(self, rep: FusedReplica, g_ptrs, w32_ptrs, w16_ptrs, flag_ptrs, aux, mc):
        from .. import ops

        ext = ops.extension()
        a = rep.arena
        n = a.padded
        mode = self.mode_override
        if mode is None:
            # measured (profiles/allreduce_update_sweep_x8.jsonl): with NVLS the two-shot kernel wins at every size
            # (14 us at 64 KB vs 42 us one-shot on 8 GPUs); without multicast one-shot only pays off for tiny
            # buckets on very small worlds, where its single barrier pair beats reading 1/N from each peer
            small = n * 4 <= self.one_shot_max_bytes and self.world <= 2 and not mc[0]
            mode = MODE_ONE_SHOT if (self.world == 1 or small) else MODE_TWO_SHOT
        use_nvls = 1 if (mc[0] and mode == MODE_TWO_SHOT and self.world > 1) else 0
        boundary = _lowp_only_boundary(rep.model, a)
        wout32 = list(w32_ptrs)
        wout16 = list(w16_ptrs) if a.w16 is not None else []
        s0 = a.slots[0].data_ptr() if len(a.slots) > 0 else 0
        s1 = a.slots[1].data_ptr() if len(a.slots) > 1 else 0
        plan = ext.FusedPlan(self.world, rep.rank, rep.opt.rule_id, mode, use_nvls, n, boundary, list(g_ptrs),
                             a.w32.data_ptr(), s0, s1, wout32, wout16, list(flag_ptrs), aux["epoch"].data_ptr(),
                             aux["step"].data_ptr(), aux["done"].data_ptr(), aux["status"].data_ptr(), mc[0], mc[1],
                             mc[2], getattr(self, "max_grid", MAX_GRID), rep.dev.cuda_index)
        plan.set_hyper(_hyper(rep.opt), 1.0)
        self.plans[id(rep)] = (plan, aux, mode, use_nvls)
        rep.comm_info = dict(mode="one_shot" if mode == MODE_ONE_SHOT else "two_shot", nvls=bool(use_nvls),
                             grid=plan.grid, fp32_bcast_from=boundary, bucket_elems=n)

    # ------------------------------------------------------------------ runtime
    def allreduce_update(self, rep: FusedReplica) -> None:
        plan, aux, _, _ = self.plans[id(rep)]
        bar = getattr(self, "_shared_gpu", None)
        if bar is not None and not torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream(rep.dev.torch_device).synchronize()
            bar.wait(timeout=120)
        plan.launch()  # on the replica's current stream; graph-capturable

    def set_opt(self, rep: FusedReplica, opt: Rule) -> None:
        plan = self.plans[id(rep)][0]
        plan.set_hyper(_hyper(opt), 1.0)

    def status(self, rep: FusedReplica):
        """(nonfinite_grad_seen, barrier_timeout) — reads and clears the device flags."""
        aux = self.plans[id(rep)][1]
        st = aux["status"].tolist()
        aux["status"].zero_()
        return bool(st[0]), bool(st[1])

    def step_count(self, rep: FusedReplica) -> int:
        return int(self.plans[id(rep)][1]["step"].item())

    def _two_shot(self, rep) -> bool:
        return self.plans[id(rep)][2] == MODE_TWO_SHOT and self.world > 1

    def gather_masters(self, replicas) -> None:
        """Two-shot keeps fp32 masters authoritative on the owner slice only; pull the other
        slices so ``module.parameters()`` are current everywhere (checkpoints, ``train`` return)."""
        self._gather(replicas, lambda a: [a.w32])

    def gather_slots(self, replicas) -> None:
        self._gather(replicas, lambda a: list(a.slots))

    def _gather(self, replicas, pick) -> None:
        from .. import ops

        ext = ops.extension()
        reps = list(replicas)
        if not reps or not self._two_shot(reps[0]):
            for r in reps:
                r.step_count = self.step_count(r)
            return
        if len(reps) > 1:  # thread mode: peers are addressable
            for r in reps:
                torch.cuda.synchronize(r.dev.torch_device)
            nb = len(pick(reps[0].arena))
            for k in range(nb):
                for r in reps:
                    with device_ctx(r.dev):
                        ptrs = torch.tensor([pick(x.arena)[k].data_ptr() for x in reps], dtype=torch.int64,
                                            device=r.dev.torch_device)
                        ext.gather_shards(pick(r.arena)[k], ptrs, self.world, r.rank)
                for r in reps:
                    torch.cuda.synchronize(r.dev.torch_device)
        else:  # process mode: all-gather owner slices through the process group
            import torch.distributed as dist

            r = reps[0]
            for buf in pick(r.arena):
                sl = buf.numel() // self.world
                mine = buf[r.rank * sl:(r.rank + 1) * sl].clone()
                parts = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(parts, mine, group=getattr(self, "group", None))
                for i, p in enumerate(parts):
                    buf[i * sl:(i + 1) * sl].copy_(p)
        for r in reps:
            r.step_count = self.step_count(r)

    def close(self):
        self.plans.clear()
        self._keep.clear()
