"""Process-based data parallelism: one OS process per device and a gradient server.

Reference ``src/sync.jl`` (dormant there, live here): ``syncgrads`` (``:36-81``),
``getgrads`` (``:83-170``), ``start`` (``:172-232``); driver ``bin/driver.jl``.

Protocol (kept): every worker ``put``s its gradient tree into its own capacity-1
input channel, the server waits until *all* inputs are ready, takes them, reduces
to the mean and ``put``s the result into every output channel; a worker ``take``s
it and runs the optimiser locally.  All workers sending ``None`` stops the server.

Fixes: Q1 — divides by the number of contributing workers, not
``4f0``; Q3 — sleeping poll instead of a hot spin; Q11 — checkpoint directory is
created and the model name is a parameter; Q12 — ``workers`` is honoured; workers
*do* send the abort signal when they finish.

``transport="channel"`` moves gradient bytes through the channels exactly like the
reference (CPU-serialised; works without a GPU).  ``transport="fused"`` is the
B200 path: the workers form a NCCL/symmetric-memory group, gradients never leave
the GPUs — the fused kernel reduces and updates over NVLink — and the channels
only carry a per-step token (step id + NaN flag), so ``syncgrads`` degenerates to a
coordinator.
"""
from __future__ import annotations

import copy
import os
import threading
import time
import traceback
from concurrent.futures import Future
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch
from torch import nn

from ..devices import as_device, device_ctx, functional
from ..losses import logitcrossentropy
from ..optim import ADAM, Rule, state as opt_state
from ..tree import Tree, _accum, _dodiv, fmap, gradient
from ..utils.checkpoint import cycle_checkpoint_name, save_checkpoint
from ..utils.loggers import get_logger
from .channels import RemoteChannel, wait_all_ready


# --------------------------------------------------------------------------- server (C19)
def syncgrads(ip: Sequence, op: Sequence, *, verbose: bool = False, device=None,
              stop: Optional[threading.Event] = None, timeout: Optional[float] = None,
              logger=None) -> int:
    """Monitor all input channels; when every worker has published, reduce the
    gradient trees to their mean and publish the result to every output channel.

    Returns the number of synchronisation rounds served.  ``device``: where the
    reduction runs (default: CUDA device 0 if available, else CPU) — the
    reference reduces on pid 1's GPU (``sync.jl:47``).
    """
    lg = logger or get_logger()
    lg.info("Starting Syncing workers!")
    tdev = as_device(device).torch_device if device is not None else (
        torch.device("cuda", 0) if functional() else torch.device("cpu"))
    rounds = 0
    while True:
        if not wait_all_ready(ip, stop=stop, timeout=timeout):
            break
        verbose and lg.info("All workers ready")
        grads = [c.take() for c in ip]
        verbose and lg.info("take! from all workers done")
        if all(g is None for g in grads):
            lg.info("Got abort signal from every worker")
            break
        live = [g for g in grads if g is not None]
        if live and all(isinstance(g, dict) and g.get("__token__") for g in live):
            # fused transport: tokens only; agree on step + OR the NaN flags
            final = {"__token__": True, "step": max(g["step"] for g in live),
                     "nan": any(g.get("nan", False) for g in live), "n": len(live)}
        else:
            vals = [fmap(lambda t: t.to(tdev) if isinstance(t, torch.Tensor) else t, g) for g in live]
            verbose and lg.info("Moved all grads to the reduction device")
            final = vals[0]
            for v in vals[1:]:
                final = fmap(_accum, final, v)
            n = float(len(vals))  # Q1: the real worker count
            final = fmap(lambda x: _dodiv(x, n), final)
            final = fmap(lambda t: t.cpu() if isinstance(t, torch.Tensor) else t, final)
        verbose and lg.info("All grads synced")
        for g, o in zip(grads, op):
            if g is not None:  # a finished worker no longer listens
                o.put(final)
        rounds += 1
        verbose and lg.info("Set the updated_grads")
    return rounds


# --------------------------------------------------------------------------- worker (C20)
class _TokenPump:
    """Posts per-step tokens to the gradient server (``c.put`` / ``op.take``) from a helper thread.  With the fused
    transport the bytes move inside the kernel over NVLink and NaN agreement happens there too, so the server only has
    to *see* the steps (protocol parity with ``src/sync.jl:36-81``); nothing on the training path waits for it."""

    def __init__(self, c, op, lg, p):
        import queue

        self.c, self.op, self.lg, self.p = c, op, lg, p
        self.q: "queue.Queue" = queue.Queue()
        self.err: Optional[BaseException] = None
        self.served = 0
        self.t = threading.Thread(target=self._run, daemon=True, name=f"fluxdist-tokens-{p}")
        self.t.start()

    def _run(self):
        try:
            while True:
                tok = self.q.get()
                if tok is None:
                    return
                self.c.put(tok)
                self.op.take()
                self.served += 1
        except BaseException as e:  # noqa: BLE001
            self.err = e

    def post(self, tok) -> None:
        if self.err is not None:
            raise self.err
        self.q.put(tok)

    def close(self, timeout: float = 120.0) -> None:
        self.q.put(None)
        self.t.join(timeout)
        if self.err is not None:
            raise self.err


def _maybetuple(y):
    return y if isinstance(y, tuple) else (y,)


def _worker_loop(rank: int, nworkers: int, loss, tree, m: nn.Module, key, c, op, *, device, verbose: bool,
                 nsamples: int, batchsize: int, opt: Rule, class_idx, cycles: int, saveweights: bool,
                 st, val, minibatch_fn, weights_root: str, model_name: str, save_every: int,
                 transport: str, logger=None, eval_every: int = 10, use_graph: Optional[bool] = None,
                 on_step: Optional[Callable] = None, status_every: int = 10):
    from ..data import DataLoader, StagedLoader

    lg = logger or get_logger()
    dev = as_device(device)
    p = rank
    with device_ctx(dev):
        if dev.is_cuda:
            torch.cuda.set_device(dev.torch_device)
        # deep copy: tensors sent to a spawned process live in shared memory, so training the
        # received module in place would silently mutate the parent's model
        gm = copy.deepcopy(m).to(dev.torch_device)
        if val is None:
            lg.info(f"Worker {p} does not have validation set")
            val_ = minibatch_fn(tree, key, 100) if (minibatch_fn and eval_every) else None
        else:
            val_ = val
        gval = None
        if val_ is not None:
            gval = tuple(t.to(dev.torch_device) for t in val_)
            with torch.no_grad():
                vl = float(loss(gm(gval[0]), gval[1]))
            lg.info(f"Worker {p} ready to train")
            lg.info(f"Worker {p} starting with loss: {vl}")
        sts_given = st is not None
        if st is None:
            st = opt_state(opt, gm)
        else:
            st = fmap(lambda t: t.to(dev.torch_device) if isinstance(t, torch.Tensor) else t, st)

        engine = None
        tokens = None
        if transport == "fused":
            from .fused import FusedGroup

            if dev.is_cuda and any(pp.dim() == 4 for pp in gm.parameters()):
                gm = gm.to(memory_format=torch.channels_last)
            engine = FusedGroup.for_process(gm, dev, opt, loss=loss)
            rep = engine.replicas[0]
            if sts_given:
                # resume: moments and step counter go into the arena (the logical state the reference passes as `sts=`,
                # src/sync.jl:101,127-129); ADVICE r1: the fused transport used to ignore it
                rep.load_state_tree(st)
            if os.environ.get("FLUXDIST_DEBUG_EXCHANGE", "0") == "1":
                # debug mode: every worker poisons its buckets and checks the exchange (collective)
                vr = engine.verify_exchange()
                lg.info("verify_exchange", **{k: v for k, v in vr.items() if k != "errors"})
            graphed = dev.is_cuda and engine.comm.name == "native" if use_graph is None else bool(use_graph)
            # the channels carry one token per step (step id, NaN flag): posted from a helper thread so the gradient
            # server stays informed without the training loop ever waiting for a host round trip
            tokens = _TokenPump(c, op, lg, p)

        step = 0
        for n in range(1, cycles + 1):
            if n % 5 == 0:
                lg.info(f"Cycle: {n}")
            data = minibatch_fn(tree, key, nsamples)
            if engine is not None and dev.is_cuda:
                # pinned staging ring: fixed pinned slots, H2D on a copy stream, event hand-off
                nb = data[0].shape[0] // batchsize
                dl = StagedLoader.from_batches(
                    lambda i, _d=data: tuple(t[i * batchsize:(i + 1) * batchsize] for t in _d), list(range(nb)),
                    dev.torch_device) if nb > 0 else []
            else:
                dl = DataLoader(lambda b: tuple(t.to(dev.torch_device, non_blocking=True) for t in b), data,
                                batchsize=batchsize)
            for i, d in enumerate(dl, start=1):
                x, y = d
                step += 1
                if engine is not None:
                    rep = engine.replicas[0]
                    if graphed:
                        from .engine import graphed_replica_step

                        l = graphed_replica_step(rep, x, y, loss)  # one CUDA graph replay: fwd + bwd + fused exchange/update
                    else:
                        l = rep.step(x, y, loss)
                    nan = False
                    if step % status_every == 0:
                        stt = engine.check_status()  # raises on a barrier timeout (a dead peer)
                        nan = stt["nonfinite"]
                        if nan:
                            lg.info(f"Worker {p}: non-finite reduced gradient reported at step {step}")
                    tokens.post({"__token__": True, "step": step, "nan": bool(nan)})
                    if on_step is not None:
                        on_step(step, {"loss": l})
                else:
                    dm, _ = gradient(lambda mm: loss(mm(x), *_maybetuple(y)), gm)
                    verbose and lg.info(f"Worker {p} got grads")
                    c.put(fmap(lambda t: t.cpu() if isinstance(t, torch.Tensor) else t, dm))
                    verbose and lg.info(f"Worker {p} put grads in channel")
                    gnew = op.take()
                    gnew = fmap(lambda t: t.to(dev.torch_device) if isinstance(t, torch.Tensor) else t, gnew)
                    verbose and lg.info(f"Worker {p} got new grads")
                    gm, st = opt(gm, gnew, st)
                    if on_step is not None:
                        on_step(step, {})
                if eval_every and i % eval_every == 0 and gval is not None:
                    if engine is not None:
                        engine.sync_to_modules()  # fp32 masters are sharded by owner under the two-shot kernel
                    with torch.no_grad():
                        lg.info(f"Worker {p} at {i} with loss: {float(loss(gm(gval[0]), gval[1]))}")
            if saveweights and n % save_every == 0:
                if engine is not None:  # collective: every worker gathers; rank 0 writes
                    engine.sync_to_modules()
                    st = engine.state_trees()[dev]
            if saveweights and n % save_every == 0 and rank == 0:
                path = cycle_checkpoint_name(weights_root, p, model_name, n)
                save_checkpoint(path, gm, st, step=step)
                lg.info(f"Worker {p} saved {path}")
        if engine is not None:
            tokens.close()
            engine.check_status()
            engine.sync_to_modules()
            st = engine.state_trees()[dev]
            engine.close()
        if gval is not None:
            with torch.no_grad():
                lg.info(f"Worker {p} finishing with loss: {float(loss(gm(gval[0]), gval[1]))}")
        lg.info(f"Worker {p} finished - signalling")
        c.put(None)  # abort protocol (commented out in the reference, sync.jl:165)
        cpu_m = copy.deepcopy(gm).to("cpu")
        cpu_st = fmap(lambda t: t.cpu() if isinstance(t, torch.Tensor) else t, st)
        return cpu_m, cpu_st


def _proc_entry(rank, nworkers, result_q, loss, tree, m, key, c, op, kw, dist_env):
    try:
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // max(nworkers, 1)))
        if dist_env is not None:
            os.environ.update({k: str(v) for k, v in dist_env.items()})
            os.environ["RANK"] = str(rank)
            os.environ["LOCAL_RANK"] = str(rank)
            os.environ["WORLD_SIZE"] = str(nworkers)
        out = _worker_loop(rank, nworkers, loss, tree, m, key, c, op, **kw)
        from .channels import _dumps

        result_q.put((rank, "ok", _dumps(out)))
    except BaseException as e:  # noqa: BLE001
        result_q.put((rank, "err", f"{e!r}\n{traceback.format_exc()}"))
        try:
            c.put(None, timeout=5)
        except Exception:
            pass


def default_minibatch_fn(class_idx=None, onehot: bool = True):
    """``minibatch(dt, key; nsamples, class_idx)`` bound for worker use."""

    def fn(tree, key, nsamples):
        from ..data import SyntheticKey, minibatch, synthetic_minibatch

        if isinstance(key, SyntheticKey):
            return synthetic_minibatch(key, nsamples, onehot=onehot)
        if isinstance(key, (tuple, list)) and isinstance(key[0], torch.Tensor):
            ix = torch.randint(0, key[0].shape[0], (nsamples,))
            return tuple(k[ix] for k in key)
        return minibatch(tree, key, nsamples=nsamples, class_idx=class_idx, onehot=onehot)

    return fn


class _SyntheticPicklableMB:
    """Module-level callable so it can cross the spawn boundary."""

    def __init__(self, class_idx=None, onehot=True):
        self.class_idx, self.onehot = class_idx, onehot

    def __call__(self, tree, key, nsamples):
        return default_minibatch_fn(self.class_idx, self.onehot)(tree, key, nsamples)


def getgrads(loss, data_tree, ms, key, rc: Sequence, opc: Sequence, *, workers: Optional[Sequence[int]] = None,
             devices: Optional[Sequence] = None, verbose: bool = False, nsamples: int = 2000,
             batchsize: int = 48, opt: Optional[Rule] = None, class_idx=None, cycles: int = 1,
             saveweights: bool = False, sts: Optional[Sequence] = None, vals: Optional[Sequence] = None,
             minibatch_fn: Optional[Callable] = None, spawn: bool = True, weights_root: str = "weights",
             model_name: str = "model", save_every: int = 20, transport: str = "channel",
             master_port: Optional[int] = None, world: Optional[int] = None, eval_every: int = 10,
             use_graph: Optional[bool] = None, on_step: Optional[Callable] = None) -> List[Future]:
    """Launch one worker per (model, state, val, worker, device, in-chan, out-chan)
    tuple (``sync.jl:90-170``) and return futures resolving to ``(cpu_model, cpu_state)``.

    ``spawn=True`` runs each worker in its own OS process (the reference's
    ``remotecall_wait``); ``spawn=False`` runs them as threads of this process
    (useful on fake devices and in tests).

    ``workers`` may name a subset of the ``world`` workers: only those are hosted by this call.  That is how an
    externally launched job (``torchrun``: one process per GPU already exists) uses the same front door — every rank
    calls ``start(..., workers=[rank], world=N, spawn=False, transport="fused")`` with channels that reach the server
    (``StoreChannel``), and the fused transport forms its group from the job's ``torch.distributed`` world.
    """
    nw = len(rc)
    workers = list(workers) if workers is not None else list(range(nw))
    if world is not None and len(workers) != len(rc):
        # channels are indexed by worker id; keep only those of the workers hosted here
        rc = [rc[w] for w in workers]
        opc = [opc[w] for w in workers]
        if isinstance(ms, (list, tuple)) and len(ms) != len(workers):
            ms = [ms[w] for w in workers]
        nw = len(workers)
    world = world if world is not None else nw
    if not isinstance(ms, (list, tuple)):
        ms = [ms for _ in workers]  # sync.jl:83-84
    if devices is None:
        from ..devices import cuda_devices, fake_devices

        devices = cuda_devices(nw) if functional() and torch.cuda.device_count() >= nw else fake_devices(nw)
    opt = opt if opt is not None else ADAM()
    sts = list(sts) if sts is not None else [None] * nw
    vals = list(vals) if vals is not None else [None] * nw
    mb = minibatch_fn or _SyntheticPicklableMB(class_idx)
    futures: List[Future] = [Future() for _ in range(nw)]
    for f in futures:
        f.set_running_or_notify_cancel()

    def kw(i):
        return dict(device=devices[i], verbose=verbose, nsamples=nsamples, batchsize=batchsize, opt=opt,
                    class_idx=class_idx, cycles=cycles, saveweights=saveweights, st=sts[i], val=vals[i],
                    minibatch_fn=mb, weights_root=weights_root, model_name=model_name, save_every=save_every,
                    transport=transport, eval_every=eval_every, use_graph=use_graph, on_step=on_step)

    if not spawn:
        def run(i):
            try:
                futures[i].set_result(_worker_loop(workers[i], world, loss, data_tree, copy.deepcopy(ms[i]), key,
                                                   rc[i], opc[i], **kw(i)))
            except BaseException as e:  # noqa: BLE001
                try:
                    rc[i].put(None, timeout=5)
                except Exception:
                    pass
                futures[i].set_exception(e)

        for i in range(nw):
            threading.Thread(target=run, args=(i,), daemon=True, name=f"fluxdist-worker-{i}").start()
        return futures

    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    result_q = ctx.Queue()
    dist_env = None
    if transport == "fused":
        dist_env = {"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": master_port or 29533}
    procs = []
    for i in range(nw):
        pr = ctx.Process(target=_proc_entry, args=(workers[i], nw, result_q, loss, data_tree, ms[i], key,
                                                   rc[i], opc[i], kw(i), dist_env), daemon=True)
        pr.start()
        procs.append(pr)

    def collect():
        from .channels import _loads

        pending = nw
        by_rank = {w: i for i, w in enumerate(workers)}
        while pending:
            try:
                rank, status, payload = result_q.get(timeout=1.0)
            except Exception:
                if all(not p.is_alive() for p in procs) and result_q.empty():
                    for f in futures:
                        if not f.done():
                            f.set_exception(RuntimeError("worker exited without a result"))
                    return
                continue
            i = by_rank[rank]
            if status == "ok":
                futures[i].set_result(_loads(payload))
            else:
                futures[i].set_exception(RuntimeError(f"worker {rank} failed: {payload}"))
            pending -= 1
        for p in procs:
            p.join(timeout=10)

    threading.Thread(target=collect, daemon=True, name="fluxdist-collect").start()
    return futures


# --------------------------------------------------------------------------- front door (C21)
def start(loss, data_tree, key, model=None, workers: Optional[Sequence[int]] = None, rcs: Sequence = (),
          updated_grads_channel: Sequence = (), *, class_idx=None, verbose: bool = False, o=ADAM,
          opt_args: tuple = (), opt: Optional[Rule] = None, **kwargs) -> List[Future]:
    """``start(loss, data_tree, key, model, workers, input_channels, output_channels; …)``
    (``sync.jl:214-232``).  ``model`` defaults to the ResNet-50 trunk + ``Dense(1000, 200)``
    of the reference.  Forwards every keyword to :func:`getgrads` and returns its futures;
    ``[f.result() for f in futures]`` gives ``[(cpu_model, cpu_opt_state), …]``."""
    if model is None:
        from ..models import resnet_trunk_with_head

        model = resnet_trunk_with_head()
    opt = opt if opt is not None else o(*opt_args)
    return getgrads(loss, data_tree, model, key, rcs, updated_grads_channel, workers=workers, verbose=verbose,
                    opt=opt, class_idx=class_idx, **kwargs)


def make_channels(nworkers: int, spawn: bool = True):
    """``rcs`` / ``updated_grads_channel`` of ``bin/driver.jl:22-23``."""
    if spawn:
        import torch.multiprocessing as mp

        ctx = mp.get_context("spawn")
        return ([RemoteChannel(1, ctx, owner=p + 1) for p in range(nworkers)],
                [RemoteChannel(1, ctx, owner=0) for _ in range(nworkers)])
    from .channels import Channel

    return [Channel(1) for _ in range(nworkers)], [Channel(1) for _ in range(nworkers)]


def run_distributed(key, model, loss=logitcrossentropy, rcs=None, updated_grads_channel=None, *,
                    nworkers: int = 4, data_tree=None, class_idx=None, opt: Optional[Rule] = None,
                    spawn: bool = True, serve: bool = True, **kwargs):
    """``run_distributed`` of ``bin/driver.jl:25-41``: create channels (if not given),
    start the ``syncgrads`` server task, launch the workers, wait, and return
    ``(models, states)``."""
    if rcs is None or updated_grads_channel is None:
        rcs, updated_grads_channel = make_channels(nworkers, spawn)
    server = None
    if serve:
        server = threading.Thread(target=syncgrads, args=(rcs, updated_grads_channel),
                                  kwargs={"verbose": kwargs.get("verbose", False)}, daemon=True,
                                  name="fluxdist-syncgrads")
        server.start()
    futs = start(loss, data_tree, key, model, list(range(len(rcs))), rcs, updated_grads_channel,
                 class_idx=class_idx, opt=opt or ADAM(), spawn=spawn, **kwargs)
    res = [f.result() for f in futs]
    if server is not None:
        server.join(timeout=30)
    return [r[0] for r in res], [r[1] for r in res]
