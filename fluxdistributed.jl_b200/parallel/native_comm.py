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


def _bucket_fracs(world: int = 2) -> tuple:
    """Bucket plan: cumulative size fractions (in backward order) after which a bucket is cut, e.g.
    ``FLUXDIST_BUCKETS=0.55,0.9`` gives ResNet-34 three buckets — layer4 + fc (62 % of the bytes, ready first), layer3
    (31 %), the rest + all 1-D parameters — each launched on the communication stream from a gradient-ready hook while
    the backward pass of the earlier layers still runs.

    Default: ONE bucket at the end of the backward pass.  Measured on B200 (gpurun_out/r2ab, r2ab2; ResNet-34, batch
    96/GPU): 1 GPU 19.8k img/s with one bucket vs 19.4k with three early buckets; 2 GPUs 39.3k (4.887 ms/step, +53 us
    over the 1-GPU step) vs 38.3k (early grid 32) / 39.1k (early grid 64).  The exchange kernel moves its bytes with SM
    instructions (peer loads / multimem.ld_reduce), so running it early takes issue slots and CTA slots from the
    backward pass it overlaps with — more than the ~50-100 us of exposed exchange time it hides at these sizes."""
    env = os.environ.get("FLUXDIST_BUCKETS")
    if env is None or env.strip() in ("", "0", "off", "none"):
        return ()
    return tuple(float(x) for x in env.split(","))


def _early_grid() -> int:
    """CTAs of a bucket kernel launched from a trigger: it shares the GPU with the backward pass of the earlier layers,
    so it gets a bounded grid (the single-launch BN kernels keep the same number of CTA slots free, csrc/nn_kernels.cu)."""
    return max(1, int(os.environ.get("FLUXDIST_EARLY_GRID", "32")))


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
        if id(p) in lowp_ids and off >= boundary:  # gaps = bucket padding (no parameters live there)
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
        shared = max(cu.count(c) for c in set(cu)) > 1  # virtual replicas on one GPU: no early buckets
        for i, (m, d) in enumerate(zip(models, devs)):
            rep = FusedReplica(m, d, opt, n, i, lowp, order="backward",
                               bucket_fracs=(() if shared else _bucket_fracs(n)))
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
        aux = [comm._make_aux(rep) for rep in reps]
        for i, rep in enumerate(reps):
            comm._make_plans(rep, g=[r.arena.g32 for r in reps], w32=[r.arena.w32 for r in reps],
                             w16=[r.arena.w16 for r in reps], flags=[x["flags"] for x in aux], aux=aux[i], mc=(0, 0, 0))
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
            rep = FusedReplica(model, dev, opt, 1, 0, lowp, order="backward", bucket_fracs=_bucket_fracs(1))
            g_ptrs, w32_ptrs = [rep.arena.g32.data_ptr()], [rep.arena.w32.data_ptr()]
            w16_ptrs = [rep.arena.w16.data_ptr() if rep.arena.w16 is not None else 0]
            flags = torch.zeros(comm.FLAG_WORDS, dtype=torch.int32, device=tdev)
            flag_ptrs = [flags.data_ptr()]
        else:
            import torch.distributed._symmetric_memory as symm

            gname = (group or dist.group.WORLD).group_name

            def alloc(nel, dt):
                t = symm.empty(nel, dtype=dt, device=tdev)
                t.zero_()
                return t

            rep = FusedReplica(model, dev, opt, world, rank, lowp, alloc=alloc, order="backward",
                               bucket_fracs=_bucket_fracs(world))
            a = rep.arena
            dist.broadcast(a.w32, src=0, group=group)  # replicas start identical
            if a.w16 is not None:
                a.w16.copy_(a.w32)
            flags = symm.empty(comm.FLAG_WORDS, dtype=torch.int32, device=tdev)
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
        aux = comm._make_aux(rep, flags=flags)
        comm._make_plans_ptrs(rep, g_ptrs, w32_ptrs, w16_ptrs, flag_ptrs, aux, mc)
        comm._keep.append(aux)
        comm.replicas = [rep]
        comm.group = group
        return rep

    # ------------------------------------------------------------------ plans (one per bucket)
    MAX_BUCKETS = 4
    FLAG_WORDS = MAX_GRID * 8 * MAX_BUCKETS  # signal pads: [bucket][cta][rank]

    def _make_aux(self, rep: FusedReplica, flags: Optional[torch.Tensor] = None) -> dict:
        """Per-replica device state of the fused kernels: signal pads and epochs per bucket, step / completion counters
        per bucket, the status words and the hyper-parameter block the kernel reads at run time."""
        with device_ctx(rep.dev):
            tdev = rep.dev.torch_device
            nb = self.MAX_BUCKETS
            return dict(flags=flags if flags is not None else torch.zeros(self.FLAG_WORDS, dtype=torch.int32, device=tdev),
                        epoch=torch.zeros(MAX_GRID * nb, dtype=torch.int32, device=tdev),
                        step=torch.zeros(nb, dtype=torch.int32, device=tdev),
                        done=torch.zeros(nb, dtype=torch.int32, device=tdev),
                        status=torch.zeros(2, dtype=torch.int32, device=tdev),
                        hyper=torch.zeros(8, dtype=torch.float32, device=tdev))

    def _make_plans(self, rep: FusedReplica, g, w32, w16, flags, aux, mc):
        self._make_plans_ptrs(rep, [t.data_ptr() for t in g], [t.data_ptr() for t in w32],
                              [(t.data_ptr() if t is not None else 0) for t in w16], [t.data_ptr() for t in flags], aux, mc)

    def _make_plans_ptrs(self, rep: FusedReplica, g_ptrs, w32_ptrs, w16_ptrs, flag_ptrs, aux, mc):
        from .. import ops

        ext = ops.extension()
        a = rep.arena
        assert len(a.buckets) <= self.MAX_BUCKETS
        boundary = _lowp_only_boundary(rep.model, a)
        plans = []
        info = []
        overlap = len(a.buckets) > 1 and os.environ.get("FLUXDIST_OVERLAP", "1") != "0"
        for b, (lo, hi) in enumerate(a.buckets):
            n = hi - lo
            mode = self.mode_override
            if mode is None:
                # measured (profiles/allreduce_update_sweep_x8.jsonl): with NVLS the two-shot kernel wins at every size
                # (14 us at 64 KB vs 42 us one-shot on 8 GPUs); without multicast one-shot only pays off for tiny
                # buckets on very small worlds, where its single barrier pair beats reading 1/N from each peer
                small = n * 4 <= self.one_shot_max_bytes and self.world <= 2 and not mc[0]
                mode = MODE_ONE_SHOT if (self.world == 1 or small) else MODE_TWO_SHOT
            use_nvls = 1 if (mc[0] and mode == MODE_TWO_SHOT and self.world > 1) else 0
            early = overlap and a.bucket_triggers[b] is not None
            max_grid = min(getattr(self, "max_grid", MAX_GRID), _early_grid() if early else MAX_GRID)
            off32, off16 = lo * 4, lo * 2
            s0 = a.slots[0].data_ptr() + off32 if len(a.slots) > 0 else 0
            s1 = a.slots[1].data_ptr() + off32 if len(a.slots) > 1 else 0
            wout16 = [(p + off16 if p else 0) for p in w16_ptrs] if a.w16 is not None else []
            plan = ext.FusedPlan(self.world, rep.rank, rep.opt.rule_id, mode, use_nvls, n, max(min(boundary - lo, n), 0),
                                 [p + off32 for p in g_ptrs], a.w32.data_ptr() + off32, s0, s1,
                                 [p + off32 for p in w32_ptrs], wout16,
                                 [p + b * MAX_GRID * 8 * 4 for p in flag_ptrs],
                                 aux["epoch"].data_ptr() + b * MAX_GRID * 4, aux["step"].data_ptr() + b * 4,
                                 aux["done"].data_ptr() + b * 4, aux["status"].data_ptr(),
                                 (mc[0] + off32) if mc[0] else 0, (mc[1] + off32) if mc[1] else 0,
                                 (mc[2] + off16) if mc[2] else 0, max_grid, rep.dev.cuda_index)
            plan.set_hyper(_hyper(rep.opt), 1.0)
            plan.set_hyper_dev(aux["hyper"].data_ptr())
            plans.append(plan)
            info.append(dict(elems=n, mode="one_shot" if mode == MODE_ONE_SHOT else "two_shot", nvls=bool(use_nvls),
                             grid=plan.grid, early=bool(early)))
        self.plans[id(rep)] = (plans, aux, [i["mode"] == "two_shot" for i in info])
        self._write_hyper(rep, rep.opt)
        rep.overlap = overlap
        rep.comm_stream = None
        if overlap and rep.dev.is_cuda:
            with device_ctx(rep.dev):
                rep.comm_stream = torch.cuda.Stream(device=rep.dev.torch_device, priority=-1)
            ext.bn_set_reserve(_early_grid())  # the single-launch BN kernels leave room for a concurrent bucket kernel
        rep.comm_info = dict(mode=info[-1]["mode"], nvls=info[-1]["nvls"], grid=info[-1]["grid"],
                             fp32_bcast_from=boundary, bucket_elems=a.padded, buckets=info, overlap=bool(overlap))

    def _write_hyper(self, rep: FusedReplica, opt: Rule) -> None:
        aux = self.plans[id(rep)][1]
        h = _hyper(opt) + [0.0, 0.0, 1.0]
        with device_ctx(rep.dev):
            # stream-ordered: replays already queued keep the old values, later ones see the new ones
            aux["hyper"].copy_(torch.tensor(h[:8], dtype=torch.float32), non_blocking=False)

    # ------------------------------------------------------------------ runtime
    def launch_bucket(self, rep: FusedReplica, b: int, early: bool) -> None:
        plans = self.plans[id(rep)][0]
        bar = getattr(self, "_shared_gpu", None)
        if bar is not None and not torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream(rep.dev.torch_device).synchronize()
            bar.wait(timeout=120)
        cs = rep.comm_stream
        if early and cs is not None:
            from ..ops import native as _native

            tdev = rep.dev.torch_device
            cur = torch.cuda.current_stream(tdev)
            cs.wait_stream(cur)                     # BN-affine / bias gradients and everything queued so far
            key = tdev.index if tdev.index is not None else torch.cuda.current_device()
            if key in _native._SIDE_PENDING:        # weight gradients written by the side-stream wgrad kernels
                cs.wait_stream(_native._SIDE[key])
            with torch.cuda.stream(cs):
                plans[b].launch()
            rep._comm_pending = True
        else:
            plans[b].launch()  # on the replica's current stream; graph-capturable

    def join(self, rep: FusedReplica) -> None:
        """The step's last act on the compute stream: wait for the bucket kernels launched on the communication stream
        (the next forward reads the weights they broadcast)."""
        if getattr(rep, "_comm_pending", False) and rep.comm_stream is not None:
            torch.cuda.current_stream(rep.dev.torch_device).wait_stream(rep.comm_stream)
            rep._comm_pending = False

    def allreduce_update(self, rep: FusedReplica) -> None:
        for b in range(len(self.plans[id(rep)][0])):
            self.launch_bucket(rep, b, early=False)

    def set_opt(self, rep: FusedReplica, opt: Rule) -> None:
        """Hyper-parameters live in device memory (read by the kernel at run time), so LR schedules and ``set_opt`` take
        effect on the very next step even when the step is a captured CUDA graph."""
        for plan in self.plans[id(rep)][0]:
            plan.set_hyper(_hyper(opt), 1.0)
        self._write_hyper(rep, opt)

    def set_max_spins(self, rep: FusedReplica, spins: int) -> None:
        """Bound of the device-side barrier spin (turns a dead peer into a status flag instead of a hang)."""
        for plan in self.plans[id(rep)][0]:
            plan.set_max_spins(int(spins))

    def set_step(self, rep: FusedReplica, step: int) -> None:
        """Optimiser step counter of every bucket (Adam bias correction is derived from it on the device)."""
        aux = self.plans[id(rep)][1]
        with device_ctx(rep.dev):
            aux["step"].fill_(int(step))

    def status(self, rep: FusedReplica):
        """(nonfinite_grad_seen, barrier_timeout) — reads and clears the device flags."""
        aux = self.plans[id(rep)][1]
        st = aux["status"].tolist()
        aux["status"].zero_()
        return bool(st[0]), bool(st[1])

    def step_count(self, rep: FusedReplica) -> int:
        return int(self.plans[id(rep)][1]["step"][0].item())

    def _two_shot(self, rep) -> bool:
        return any(self.plans[id(rep)][2]) and self.world > 1

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
        two_shot = self.plans[id(reps[0])][2]
        if len(reps) > 1:  # thread mode: peers are addressable
            for r in reps:
                torch.cuda.synchronize(r.dev.torch_device)
            nb = len(pick(reps[0].arena))
            for k in range(nb):
                for r in reps:
                    with device_ctx(r.dev):
                        for bi, (lo, hi) in enumerate(r.arena.buckets):  # ownership is per bucket
                            if not two_shot[bi]:
                                continue
                            ptrs = torch.tensor([pick(x.arena)[k].data_ptr() + lo * 4 for x in reps], dtype=torch.int64,
                                                device=r.dev.torch_device)
                            ext.gather_shards(pick(r.arena)[k][lo:hi], ptrs, self.world, r.rank)
                for r in reps:
                    torch.cuda.synchronize(r.dev.torch_device)
        else:  # process mode: all-gather owner slices through the process group
            import torch.distributed as dist

            r = reps[0]
            for buf in pick(r.arena):
                for bi, (lo, hi) in enumerate(r.arena.buckets):
                    if not two_shot[bi]:
                        continue
                    sl = (hi - lo) // self.world
                    mine = buf[lo + r.rank * sl: lo + (r.rank + 1) * sl].clone()
                    parts = [torch.empty_like(mine) for _ in range(self.world)]
                    dist.all_gather(parts, mine, group=getattr(self, "group", None))
                    for i, p in enumerate(parts):
                        buf[lo + i * sl: lo + (i + 1) * sl].copy_(p)
        for r in reps:
            r.step_count = self.step_count(r)

    def close(self):
        self.plans.clear()
        self._keep.clear()
