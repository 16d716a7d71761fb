"""The fused gradient-exchange + optimiser engine (the B200 product path).

What the reference does per step (``src/ddp_tasks.jl:59-109,163-172``) in
~(N+2)·L small launches plus 2·L·N peer copies — gather grads to HOST, N-1 add
passes, a divide pass, broadcast back, per-tensor optimiser — is here ONE kernel
launch per replica over flat arenas:

    barrier-in -> reduce my slice from every peer (P2P ``ld.global`` or in-switch
    ``multimem.ld_reduce``) -> x 1/N -> Momentum/Adam in registers on the fp32
    master shard -> store new weights (bf16 and/or fp32) to every replica
    (P2P stores or ``multimem.st``) -> barrier-out

* ``FlatArena``     one allocation each for master weights, gradients, moments
                    and (optionally) bf16 compute weights; identical layout on
                    every replica so a single offset addresses all peers.
* ``FusedReplica``  a model bound to its arena: ``p.data``/``p.grad`` are views,
                    so autograd (or the native wgrad kernels) write gradients
                    straight into the arena — the reference's ``markbuffer!``
                    copies disappear.
* ``FusedGroup``    N replicas plus a communicator:
    ``NativeComm``  hand-written sm_100a kernel over peer-mapped memory
                    (threads of one process via UVA peer access, or processes
                    via torch symmetric memory + NVLS multicast),
    ``ThreadComm``  CPU / no-extension oracle for task mode,
    ``DistComm``    ``torch.distributed`` all-reduce (gloo on CPU; on CUDA this
                    is the "only calls NCCL" *baseline*, never the default).

Two-shot shards the update work and the moments by owner rank (ZeRO-1-like);
``state_trees()`` re-assembles the logical per-replica optimiser state the
reference API exposes (``sts[dev]``).
"""
from __future__ import annotations

import math
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch
from torch import nn

from ..devices import as_device, device_ctx
from ..optim import Rule, flat_step
from ..optim.rules import RULE_ADAM, RULE_ADAMW
from ..tree import Tree, _module_tree

ALIGN = 2048  # elements; keeps every rank slice 16B-aligned for v4 accesses at any N<=8 and bf16


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _sequential_units(model: nn.Module) -> List[nn.Module]:
    """Execution-ordered units of a sequentially composed model: the root's children in registration order with
    ``nn.Sequential`` containers expanded recursively.  (Registration order == execution order is the convention of
    the model zoo and of ``Chain``-style models; a model that violates it simply must not be bucketed.)"""
    out: List[nn.Module] = []

    def walk(m: nn.Module, top: bool):
        kids = list(m.children())
        if kids and (top or isinstance(m, nn.Sequential)):
            for k in kids:
                walk(k, False)
        else:
            out.append(m)

    walk(model, True)
    return out


class FlatArena:
    """Flat, identically laid out buffers for one replica.

    ``order="backward"`` lays the parameters out in the order their gradients become ready in the backward pass
    (matrices/filters of the LAST layers first, 1-D parameters at the very end) and cuts the arena into ``buckets``
    (element ranges whose ends are multiples of ``ALIGN * world``): bucket 0 (the deepest layers, most of the
    bytes) can be exchanged while the backward pass of the earlier layers still runs (SURVEY §5.8 bucket plan;
    reference ``src/ddp_tasks.jl:203-228`` has two global barriers instead).  ``bucket_triggers[b]`` is the module
    whose *input gradient* signals that every gradient of bucket ``b`` has been produced (None: end of backward).
    """

    def __init__(self, model: nn.Module, device: torch.device, n_slots: int, lowp: Optional[torch.dtype],
                 world: int, alloc: Optional[Callable[[int, torch.dtype], torch.Tensor]] = None,
                 sort_lowp_first: bool = False, order: str = "forward", bucket_fracs: Sequence[float] = ()):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        names = [n for n, _ in named]
        cuts = set()
        self.bucket_triggers: List[Optional[nn.Module]] = []
        if order == "backward":
            # units = the modules a sequential model executes one after the other (containers are expanded); the
            # gradient of a unit's INPUT is complete only after the whole unit (all its branches) ran backward, which
            # makes "input gradient of unit u" a safe signal for "every gradient of u and of all later units exists"
            units = _sequential_units(model)
            name_of = {id(p): n for n, p in named}
            seen, mats, vecs, unit_end = set(), [], [], {}
            for u in reversed(units):
                ps = [p for p in u.parameters() if p.requires_grad and id(p) not in seen]
                seen.update(id(p) for p in ps)
                for p in reversed(ps):
                    (mats if p.dim() >= 2 else vecs).append(p)
                if mats:
                    unit_end[len(mats) - 1] = u  # arena index of the unit's last matrix -> the unit
            rest = [p for p in self.params if id(p) not in seen]  # parameters outside the unit walk (none for the zoo)
            for p in reversed(rest):
                (mats if p.dim() >= 2 else vecs).append(p)
            total = sum(_round_up(p.numel(), 8) for p in mats + vecs)
            if bucket_fracs and total >= (1 << 20) and not rest:
                acc, fr = 0, sorted(bucket_fracs)
                for i, p in enumerate(mats[:-1]):
                    acc += _round_up(p.numel(), 8)
                    if fr and i in unit_end and acc >= fr[0] * total:
                        cuts.add(i)
                        self.bucket_triggers.append(unit_end[i])
                        while fr and acc >= fr[0] * total:
                            fr.pop(0)
            self.params = mats + vecs
            names = [name_of[id(p)] for p in self.params]
        elif sort_lowp_first:  # matrices/filters first, 1-D (BN affine, biases) last: see native_comm
            pairs = sorted(zip(names, self.params), key=lambda np_: np_[1].dim() < 2)
            names = [n for n, _ in pairs]
            self.params = [p for _, p in pairs]
        self.bucket_triggers.append(None)  # the last bucket goes out at the end of the backward pass
        self.names = names
        gran = ALIGN * max(world, 1)
        self.offsets: List[int] = []
        self.buckets: List[tuple] = []
        off, lo = 0, 0
        for i, p in enumerate(self.params):
            self.offsets.append(off)
            off += _round_up(p.numel(), 8)  # 16B-aligned starts for bf16 and fp32 views
            if i in cuts:
                off = _round_up(off, gran)
                self.buckets.append((lo, off))
                lo = off
        self.numel = off
        self.padded = _round_up(max(off, 1), gran)
        self.buckets.append((lo, self.padded))
        mk = alloc or (lambda n, dt: torch.zeros(n, dtype=dt, device=device))
        self.w32 = mk(self.padded, torch.float32)
        self.g32 = mk(self.padded, torch.float32)
        self.slots = [torch.zeros(self.padded, dtype=torch.float32, device=device) for _ in range(n_slots)]
        self.w16 = mk(self.padded, lowp) if lowp is not None else None
        self.device = device

    def view(self, flat: torch.Tensor, i: int) -> torch.Tensor:
        p = self.params[i]
        v = flat[self.offsets[i]: self.offsets[i] + p.numel()]
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            k, c, r, s = p.shape
            return v.view(k, r, s, c).permute(0, 3, 1, 2)  # physical KRSC, logical KCRS
        return v.view(p.shape)

    def owner_ranges(self, world: int):
        """[(lo, hi, rank)] element ranges whose fp32 master / moments are authoritative on ``rank`` under the
        two-shot (sharded) update: every bucket is split evenly over the ranks."""
        out = []
        for lo, hi in self.buckets:
            sl = (hi - lo) // world
            for r in range(world):
                out.append((lo + r * sl, lo + (r + 1) * sl, r))
        return out


class FusedReplica:
    def __init__(self, model: nn.Module, dev, opt: Rule, world: int, rank: int, lowp: Optional[torch.dtype],
                 alloc=None, sort_lowp_first: bool = False, order: str = "forward", bucket_fracs: Sequence[float] = ()):
        self.model, self.dev, self.opt = model, as_device(dev), opt
        self.world, self.rank = world, rank
        self.step_count = 0
        self.group: Optional["FusedGroup"] = None
        self.stream = None  # set by the group when several replicas share one GPU (virtual replicas)
        self._armed = False          # inside forward_backward: bucket triggers may fire
        self._launched: set = set()  # buckets already handed to the communicator this step
        tdev = self.dev.torch_device
        with device_ctx(self.dev):
            self.arena = FlatArena(model, tdev, opt.n_slots, lowp, world, alloc, sort_lowp_first, order, bucket_fracs)
            a = self.arena
            for b, trig in enumerate(a.bucket_triggers):
                if trig is not None:
                    trig.register_forward_pre_hook(self._make_trigger(b))
            with torch.no_grad():
                for i, p in enumerate(a.params):
                    a.view(a.w32, i).copy_(p.detach().float())
                    p.data = a.view(a.w32, i)       # the module now *is* the arena
                    p.grad = a.view(a.g32, i)       # autograd accumulates in place
                    p._gradbuf = p.grad             # native wgrad kernels write here
                    if a.w16 is not None:
                        p._lowp = a.view(a.w16, i)
                if a.w16 is not None:
                    a.w16.copy_(a.w32)

    # ---- bucket triggers: the gradient of a trigger module's input is complete only after every later layer ran backward
    def _make_trigger(self, b: int):
        def pre_hook(_mod, args):
            x = args[0] if args else None
            if self._armed and isinstance(x, torch.Tensor) and x.requires_grad:
                x.register_hook(lambda g, _b=b: self._bucket_ready(_b))
        return pre_hook

    def _bucket_ready(self, b: int) -> None:
        comm = self.group.comm if self.group is not None else None
        if not self._armed or comm is None or not hasattr(comm, "launch_bucket") or b in self._launched:
            return None
        # everything up to and including b is complete (buckets are in backward order)
        for k in range(b + 1):
            if k not in self._launched:
                self._launched.add(k)
                comm.launch_bucket(self, k, early=True)
        return None

    # ---- one training step on this replica
    def forward_backward(self, x, y, loss: Callable) -> torch.Tensor:
        with device_ctx(self.dev):
            self._launched = set()
            self._armed = bool(getattr(self, "overlap", False))
            try:
                return self._forward_backward(x, y, loss)
            finally:
                self._armed = False

    def _forward_backward(self, x, y, loss: Callable) -> torch.Tensor:
        with device_ctx(self.dev):
            self.arena.g32.zero_()
            for i, p in enumerate(self.arena.params):  # hooks may have replaced .grad
                if p.grad is None or p.grad.data_ptr() != self.arena.view(self.arena.g32, i).data_ptr():
                    p.grad = self.arena.view(self.arena.g32, i)
            amp = self.arena.w16 is not None and self.dev.is_cuda
            if amp and x.dtype == torch.float32:
                x = x.to(self.arena.w16.dtype)
            if amp:
                with torch.autocast("cuda", dtype=self.arena.w16.dtype):
                    out = self.model(x)
                    l = loss(out, y)
            else:
                out = self.model(x)
                l = loss(out, y)
            l.backward()
            return l.detach()

    def reduce_and_update(self) -> None:
        """Exchange + update of every bucket not yet launched by a trigger, then join the communication stream."""
        assert self.group is not None
        if self.dev.is_cuda:
            from ..ops import native as _native

            _native.join_side_stream(self.dev.torch_device)  # weight gradients queued on the side stream
        self.step_count += 1
        comm = self.group.comm
        if hasattr(comm, "launch_bucket"):
            for b in range(len(self.arena.buckets)):
                if b not in self._launched:
                    comm.launch_bucket(self, b, early=False)
            comm.join(self)
            self._launched = set()
        else:
            comm.allreduce_update(self)

    def zero_grads(self) -> None:
        self.arena.g32.zero_()

    # ---- snapshot / restore of everything a training step mutates (used to make graph warm-up side-effect free)
    def snapshot_state(self) -> dict:
        a = self.arena
        with device_ctx(self.dev):
            snap = {"w32": a.w32.clone(), "slots": [s.clone() for s in a.slots],
                    "w16": a.w16.clone() if a.w16 is not None else None,
                    "buffers": [b.clone() for b in self.model.buffers()], "step_count": self.step_count}
            comm = self.group.comm if self.group is not None else None
            if comm is not None and hasattr(comm, "plans") and id(self) in getattr(comm, "plans", {}):
                snap["dev_step"] = comm.plans[id(self)][1]["step"].clone()
                # the sticky status words too: a non-finite gradient seen during a graph warm-up on the (zero-filled)
                # static buffers must not surface later as a training event
                snap["dev_status"] = comm.plans[id(self)][1]["status"].clone()
        return snap

    def restore_state(self, snap: dict) -> None:
        a = self.arena
        with device_ctx(self.dev), torch.no_grad():
            a.w32.copy_(snap["w32"])
            for s, t in zip(a.slots, snap["slots"]):
                s.copy_(t)
            if a.w16 is not None and snap["w16"] is not None:
                a.w16.copy_(snap["w16"])
            for b, t in zip(self.model.buffers(), snap["buffers"]):
                b.copy_(t)
            self.step_count = snap["step_count"]
            if "dev_step" in snap:
                self.group.comm.plans[id(self)][1]["step"].copy_(snap["dev_step"])
            if "dev_status" in snap:
                self.group.comm.plans[id(self)][1]["status"].copy_(snap["dev_status"])

    def step(self, x, y, loss: Callable, fused_update: bool = True):
        if self.stream is not None:
            with device_ctx(self.dev), torch.cuda.stream(self.stream):
                self.stream.wait_stream(torch.cuda.default_stream(self.dev.torch_device))
                l = self.forward_backward(x, y, loss)
                if fused_update:
                    self.reduce_and_update()
                return l
        l = self.forward_backward(x, y, loss)
        if fused_update:
            self.reduce_and_update()
        return l

    # ---- logical optimiser state in the reference's tree form
    def state_tree(self) -> Tree:
        a = self.arena
        idx = {id(p): i for i, p in enumerate(a.params)}
        t = self.step_count

        def leaf(p):
            if id(p) not in idx:
                return None
            i = idx[id(p)]
            sl = [a.view(s, i).clone() for s in a.slots]
            if self.opt.rule_id in (RULE_ADAM, RULE_ADAMW):
                b1, b2 = self.opt.beta
                return (sl[0], sl[1], (b1 ** (t + 1), b2 ** (t + 1)))
            return sl[0] if sl else None

        return _module_tree(self.model, leaf)

    def load_state_tree(self, st: Tree, step_count: int = 0) -> None:
        from ..tree import leaves

        a = self.arena
        vals = [v for _, v in leaves(_module_tree(self.model, lambda p: p))]
        flat = [v for _, v in leaves(st)] if st is not None else []
        if not step_count:
            # Adam-family states carry the running beta powers (beta^(t+1)): recover the step from them so the
            # kernel's bias correction continues where the saved state left off (resume restarted it)
            for v in flat:
                if isinstance(v, tuple) and len(v) == 3 and isinstance(v[2], tuple) and self.opt.rule_id in (RULE_ADAM, RULE_ADAMW):
                    b1 = float(self.opt.beta[0])
                    p1 = float(v[2][0])
                    if 0.0 < p1 < 1.0 and 0.0 < b1 < 1.0:
                        step_count = max(int(round(math.log(p1) / math.log(b1))) - 1, 0)
                    break
        self.step_count = step_count
        comm = self.group.comm if self.group is not None else None
        if comm is not None and hasattr(comm, "set_step"):
            comm.set_step(self, step_count)
        # walk model params in order and copy moments in
        idx = {id(p): i for i, p in enumerate(a.params)}
        model_params = [p for p in self.model.parameters() if p.requires_grad]
        for p, s in zip(model_params, self._state_leaves(st)):
            i = idx[id(p)]
            if s is None:
                continue
            if isinstance(s, tuple):
                a.view(a.slots[0], i).copy_(s[0])
                a.view(a.slots[1], i).copy_(s[1])
            else:
                a.view(a.slots[0], i).copy_(s)

    def _state_leaves(self, st):
        def rec(m, s):
            for name, p in m.named_parameters(recurse=False):
                if p.requires_grad:
                    yield None if s is None else s.get(name)
            for name, c in m.named_children():
                yield from rec(c, None if s is None else s.get(name))

        return rec(self.model, st)


# =========================================================================== communicators
class ThreadComm:
    """Oracle communicator for threads of one process (CPU, or CUDA without the
    extension): barrier, every replica averages all gradient arenas in rank order
    (bit-identical on every replica), applies ``flat_step`` to its full arena."""

    name = "thread"

    def __init__(self, replicas: Sequence[FusedReplica], timeout: float = 300.0):
        self.replicas = list(replicas)
        n = len(self.replicas)
        self.timeout = timeout
        self._b1 = threading.Barrier(n) if n > 1 else None
        self._b2 = threading.Barrier(n) if n > 1 else None

    def abort(self) -> None:
        """Break the barriers so that replicas blocked in them raise instead of waiting for a peer that failed."""
        for b in (self._b1, self._b2):
            if b is not None:
                b.abort()

    def reset(self) -> None:
        for b in (self._b1, self._b2):
            if b is not None:
                b.reset()

    def allreduce_update(self, rep: FusedReplica) -> None:
        if self._b1 is not None:
            if rep.dev.is_cuda:
                torch.cuda.synchronize(rep.dev.torch_device)
            self._b1.wait(timeout=self.timeout)
        a = rep.arena
        acc = torch.zeros_like(a.g32)
        for r in self.replicas:  # fixed order -> identical rounding everywhere
            acc += r.arena.g32.to(a.device)
        acc /= float(len(self.replicas))
        flat_step(rep.opt, a.w32, acc, a.slots, rep.step_count)
        if a.w16 is not None:
            a.w16.copy_(a.w32)
        if self._b2 is not None:
            if rep.dev.is_cuda:
                torch.cuda.synchronize(rep.dev.torch_device)
            self._b2.wait(timeout=self.timeout)

    def close(self):
        pass


class DistComm:
    """``torch.distributed`` all-reduce + flat update (gloo on CPU; NCCL baseline on CUDA)."""

    name = "dist"

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)

    def allreduce_update(self, rep: FusedReplica) -> None:
        a = rep.arena
        if self.world > 1:
            self.dist.all_reduce(a.g32, op=self.dist.ReduceOp.SUM, group=self.group)
            a.g32.div_(float(self.world))
        flat_step(rep.opt, a.w32, a.g32, a.slots, rep.step_count)
        if a.w16 is not None:
            a.w16.copy_(a.w32)

    def close(self):
        pass


class FusedGroup:
    def __init__(self, replicas: Sequence[FusedReplica], comm, loss=None):
        self.replicas = list(replicas)
        self.comm = comm
        self.loss = loss
        # task mode replays one CUDA graph per replica and step when the native kernel path is active
        self.use_graphs = getattr(comm, "name", "") == "native" and all(r.dev.is_cuda for r in self.replicas) \
            and len({r.dev.cuda_index for r in self.replicas}) == len(self.replicas)
        for r in self.replicas:
            r.group = self
        # replicas that share a physical GPU must run on distinct streams, otherwise the device-side
        # barrier of the fused kernel would wait for a kernel queued behind it on the same stream
        cu = [r.dev.cuda_index for r in self.replicas if r.dev.is_cuda]
        if len(cu) != len(set(cu)):
            from ..ops import native as _native

            _native.BN_FUSED = False  # grid-barrier kernels of co-located replicas could starve each other
            for r in self.replicas:
                with device_ctx(r.dev):
                    r.stream = torch.cuda.Stream(device=r.dev.torch_device)

    # ---- constructors
    @staticmethod
    def _lowp(dtype, devs) -> Optional[torch.dtype]:
        if dtype is None:
            return torch.bfloat16 if all(d.is_cuda for d in devs) else None
        return None if dtype == torch.float32 else dtype

    @classmethod
    def for_threads(cls, models: Sequence[nn.Module], devs: Sequence, opt: Rule, loss=None, dtype=None,
                    comm: str = "auto", **kw) -> "FusedGroup":
        """Task mode: N replicas in this process."""
        devs = [as_device(d) for d in devs]
        lowp = cls._lowp(dtype, devs)
        n = len(models)
        use_native = comm == "native" or (comm == "auto" and all(d.is_cuda for d in devs) and _native_ok())
        alloc = None
        if use_native:
            from .native_comm import NativeComm

            reps = NativeComm.build_thread_replicas(models, devs, opt, lowp, **kw)
            return cls(reps, reps[0]._native_comm, loss)
        reps = [FusedReplica(m, d, opt, n, i, lowp, alloc) for i, (m, d) in enumerate(zip(models, devs))]
        return cls(reps, ThreadComm(reps), loss)

    @classmethod
    def for_process(cls, model: nn.Module, dev, opt: Rule, loss=None, dtype=None, comm: str = "auto",
                    group=None, **kw) -> "FusedGroup":
        """Process mode: this process owns one replica; peers are other ranks of the
        default (or given) ``torch.distributed`` group."""
        import torch.distributed as dist

        import os

        dev = as_device(dev)
        if not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            backend = "nccl" if dev.is_cuda else "gloo"
            dist.init_process_group(backend, device_id=dev.torch_device if dev.is_cuda else None)
        if dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        else:  # a single stand-alone process: world of one, no process group needed
            world, rank = 1, 0
        lowp = cls._lowp(dtype, [dev])
        use_native = comm == "native" or (comm == "auto" and dev.is_cuda and _native_ok())
        if use_native:
            from .native_comm import NativeComm

            rep = NativeComm.build_process_replica(model, dev, opt, lowp, world, rank, group, **kw)
            return cls([rep], rep._native_comm, loss)
        rep = FusedReplica(model, dev, opt, world, rank, lowp)
        # replicas must start identical (the reference copies one host model to every device)
        if world > 1:
            dist.broadcast(rep.arena.w32, src=0, group=group)
            if rep.arena.w16 is not None:
                rep.arena.w16.copy_(rep.arena.w32)
            return cls([rep], DistComm(group), loss)
        return cls([rep], ThreadComm([rep]), loss)

    # ---- API used by task.train / process workers
    def set_opt(self, opt: Rule) -> None:
        for r in self.replicas:
            r.opt = opt
            if hasattr(self.comm, "set_opt"):
                self.comm.set_opt(r, opt)

    def check_status(self, raise_on_timeout: bool = True) -> Dict[str, bool]:
        """Read (and clear) the device-side status words of the fused kernels: ``nonfinite`` — a reduced gradient was
        NaN/Inf (reference ``check_nans``, ``src/ddp_tasks.jl:86-91``, done in-kernel); ``barrier_timeout`` — a
        cross-GPU barrier gave up waiting for a peer, in which case the kernel returned WITHOUT updating and the
        replicas can no longer be assumed in sync: raise.  Costs one tiny D2H copy per replica, so the
        training loops call it every few steps, not every step."""
        out = {"nonfinite": False, "barrier_timeout": False}
        if hasattr(self.comm, "status"):
            for r in self.replicas:
                nan, tmo = self.comm.status(r)
                out["nonfinite"] |= bool(nan)
                out["barrier_timeout"] |= bool(tmo)
        if out["barrier_timeout"] and raise_on_timeout:
            raise RuntimeError("fused all-reduce+update: a cross-GPU barrier timed out (a peer replica died or stalled); "
                               "the step was not applied on at least one replica — replicas are out of sync")
        return out

    def skip_step(self) -> None:
        """Collective skip (reference OOM handling, ``src/ddp_tasks.jl:230-237``): every replica drops its gradients
        and NO replica enters the exchange, so nobody waits in a device-side barrier for a peer that bailed out."""
        for r in self.replicas:
            with device_ctx(r.dev):
                r.arena.g32.zero_()
                r._launched = set()
                r._armed = False

    def sync_to_modules(self) -> None:
        """Make every module's fp32 parameters current (gathers owner shards when
        the communicator keeps sharded masters)."""
        if hasattr(self.comm, "gather_masters"):
            self.comm.gather_masters(self.replicas)

    def state_trees(self) -> Dict[Any, Tree]:
        if hasattr(self.comm, "gather_slots"):
            self.comm.gather_slots(self.replicas)
        return {r.dev: r.state_tree() for r in self.replicas}

    def ensure_synced(self) -> bool:
        """Replica-consistency check (reference ``ensure_synced``, ``ddp_tasks.jl:115-126``) as a cheap
        collective: compare a checksum of every replica's compute weights.  Replicas are expected to be
        bit-identical after every fused step."""
        def digest(r):
            a = r.arena
            w = a.w16 if a.w16 is not None else a.w32
            v = w.view(torch.int16 if w.element_size() == 2 else torch.int32).to(torch.int64)
            return int((v * 2654435761 % 4294967291).sum().item())

        sums = [digest(r) for r in self.replicas]
        if len(self.replicas) == 1:
            try:
                import torch.distributed as dist

                if dist.is_initialized() and dist.get_world_size() > 1:
                    t = torch.tensor([sums[0]], dtype=torch.int64, device=self.replicas[0].dev.torch_device)
                    lo, hi = t.clone(), t.clone()
                    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                    return bool(lo.item() == hi.item())
            except Exception:
                pass
        return len(set(sums)) == 1

    def verify_exchange(self, rounds: int = 2, raise_on_error: bool = True) -> Dict[str, Any]:
        """Debug mode of the gradient exchange: poison every bucket with a known per-rank pattern and the
        kernel's outputs with NaN, run the real ``reduce_and_update``, check every parameter element against the
        optimiser rule applied to the known mean and (native path) the barrier epochs / step counters / status words;
        state is restored afterwards.  Collective in process mode.  See ``parallel/debug.py``."""
        from .debug import verify_exchange as _verify

        return _verify(self, rounds=rounds, raise_on_error=raise_on_error)

    def close(self) -> None:
        self.comm.close()


def _native_ok() -> bool:
    from .. import ops

    return ops.native_available()
