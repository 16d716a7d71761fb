"""Fused-engine logic on CPU: flat arenas, thread communicator, gloo communicator.
The native kernel path is exercised by the gpu-marked tests; this checks that the
arena view/aliasing, step semantics and state re-assembly are right."""
import copy
import os
import socket
from concurrent.futures import ThreadPoolExecutor

import pytest
import torch
import torch.multiprocessing as mp

import fluxdist as fd
from fluxdist.parallel import FusedGroup
from fluxdist.tree import compare, gradient


def _batches(n, per=3):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(n * per, 3, 32, 32, generator=g)
    y = fd.onehot(torch.randint(0, 10, (n * per,), generator=g), 10)
    return x, y


@pytest.mark.parametrize("opt", [fd.Momentum(), fd.ADAM(), fd.Descent(0.05)])
def test_thread_group_matches_single_device(opt):
    torch.manual_seed(0)
    base = fd.tiny_convnet(10)
    n = 3
    devs = fd.fake_devices(n)
    grp = FusedGroup.for_threads([copy.deepcopy(base) for _ in range(n)], devs, opt, dtype=torch.float32)
    ref, st = copy.deepcopy(base), fd.state(opt, base)
    x, y = _batches(n)
    with ThreadPoolExecutor(n) as ex:
        for step in range(3):
            futs = [ex.submit(grp.replicas[i].step, x[3 * i:3 * i + 3], y[3 * i:3 * i + 3], fd.logitcrossentropy)
                    for i in range(n)]
            [f.result() for f in futs]
            g, _ = gradient(lambda m: fd.logitcrossentropy(m(x), y), ref)
            ref, st = opt(ref, g, st)
    for r in grp.replicas:
        for a, b in zip(r.model.parameters(), ref.parameters()):
            assert torch.allclose(a, b, atol=1e-5)
    for a, b in zip(grp.replicas[1].model.parameters(), grp.replicas[2].model.parameters()):
        assert torch.equal(a, b)  # replicas bit-identical
    sts = grp.state_trees()
    assert compare(sts[devs[0]], st, 1e-4, 1e-5)


def test_arena_views_alias_params_and_grads():
    m = fd.ResNet18(10)
    m = m.to(memory_format=torch.channels_last)
    grp = FusedGroup.for_threads([m], fd.fake_devices(1), fd.Momentum(), dtype=torch.float32)
    a = grp.replicas[0].arena
    assert a.numel >= sum(p.numel() for p in m.parameters())
    w = m.layer1[0].conv1.conv.weight
    assert w.data_ptr() >= a.w32.data_ptr() and w.data_ptr() < a.w32.data_ptr() + a.w32.numel() * 4
    assert w.shape == (64, 64, 3, 3) and w.stride() == (576, 1, 192, 64)  # physical KRSC
    x = torch.rand(2, 3, 64, 64)
    grp.replicas[0].forward_backward(x, torch.tensor([1, 2]), lambda o, y: torch.nn.functional.cross_entropy(o, y))
    assert float(a.g32.abs().sum()) > 0
    assert w.grad.data_ptr() == a.view(a.g32, a.params.index(w) if False else [id(p) for p in a.params].index(id(w))).data_ptr()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank + 10)  # different init per rank: the group must broadcast rank 0's weights
    m = fd.tiny_convnet(10)
    grp = FusedGroup.for_process(m, fd.FakeDevice(rank), fd.Momentum(), dtype=torch.float32)
    x, y = _batches(world)
    for _ in range(2):
        grp.replicas[0].step(x[3 * rank:3 * rank + 3], y[3 * rank:3 * rank + 3], fd.logitcrossentropy)
    q.put((rank, [p.detach().clone() for p in m.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_matches_single():
    """BASELINE config #1: tiny ConvNet, CIFAR-shaped synthetic data, CPU/gloo world_size=2."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(world))
    [p.join(30) for p in ps]
    torch.manual_seed(10)
    ref = fd.tiny_convnet(10)
    opt = fd.Momentum()
    st = fd.state(opt, ref)
    x, y = _batches(world)
    for _ in range(2):
        g, _ = gradient(lambda m: fd.logitcrossentropy(m(x), y), ref)
        ref, st = opt(ref, g, st)
    for a, b, c in zip(res[0], res[1], ref.parameters()):
        assert torch.equal(a, b)
        assert torch.allclose(a, c, atol=1e-5)


def test_ensure_synced_checksum():
    torch.manual_seed(0)
    base = fd.tiny_convnet(10)
    grp = FusedGroup.for_threads([copy.deepcopy(base) for _ in range(2)], fd.fake_devices(2), fd.Momentum(),
                                 dtype=torch.float32)
    assert grp.ensure_synced()
    grp.replicas[1].arena.w32[3] += 1.0
    assert not grp.ensure_synced()


@pytest.mark.parametrize("opt", [fd.Momentum(0.05, 0.9), fd.ADAM(1e-2)])
def test_fused_group_resume_from_state_trees(opt):
    """Checkpoint/resume of the fused engine: weights + logical optimiser state taken from a group after 2
    steps and loaded into a fresh group continue exactly like the uninterrupted run."""
    torch.manual_seed(0)
    base = fd.tiny_convnet(10)
    n = 2
    devs = fd.fake_devices(n)
    x, y = _batches(n)

    def run(grp, steps):
        with ThreadPoolExecutor(n) as ex:
            for _ in range(steps):
                futs = [ex.submit(grp.replicas[i].step, x[3 * i:3 * i + 3], y[3 * i:3 * i + 3], fd.logitcrossentropy)
                        for i in range(n)]
                [f.result() for f in futs]

    full = FusedGroup.for_threads([copy.deepcopy(base) for _ in range(n)], devs, opt, dtype=torch.float32)
    run(full, 4)
    first = FusedGroup.for_threads([copy.deepcopy(base) for _ in range(n)], devs, opt, dtype=torch.float32)
    run(first, 2)
    first.sync_to_modules()
    sts = first.state_trees()
    steps_done = first.replicas[0].step_count
    resumed_models = [copy.deepcopy(r.model) for r in first.replicas]
    second = FusedGroup.for_threads(resumed_models, devs, opt, dtype=torch.float32)
    for r, d in zip(second.replicas, devs):
        r.load_state_tree(sts[d], step_count=steps_done)
    run(second, 2)
    for a, b in zip(second.replicas[0].model.parameters(), full.replicas[0].model.parameters()):
        assert torch.allclose(a, b, atol=1e-6)
    assert compare(second.state_trees()[devs[0]], full.state_trees()[devs[0]], 1e-5, 1e-6)


@pytest.mark.parametrize("opt", [fd.Momentum(0.05, 0.9), fd.ADAM(1e-2), fd.Nesterov(0.05, 0.9), fd.Descent(0.1)])
def test_verify_exchange_passes_and_is_side_effect_free(opt):
    """Debug mode of the gradient exchange: poisoned buckets in, rule(mean) out on every replica, state
    untouched afterwards."""
    torch.manual_seed(0)
    base = fd.tiny_convnet(10)
    n = 3
    grp = FusedGroup.for_threads([copy.deepcopy(base) for _ in range(n)], fd.fake_devices(n), opt, dtype=torch.float32)
    x, y = _batches(n)
    with ThreadPoolExecutor(n) as ex:  # one real step first so moments / step counters are non-trivial
        [f.result() for f in [ex.submit(grp.replicas[i].step, x[3 * i:3 * i + 3], y[3 * i:3 * i + 3],
                                        fd.logitcrossentropy) for i in range(n)]]
    before = [(r.arena.w32.clone(), [s.clone() for s in r.arena.slots], r.step_count) for r in grp.replicas]
    rep = grp.verify_exchange(rounds=2)
    assert rep["ok"] and rep["replicas"] == n and rep["checked_elems"] > 0 and rep["max_rel_err"] < 1e-4
    for r, (w, sl, sc) in zip(grp.replicas, before):
        assert torch.equal(r.arena.w32, w) and r.step_count == sc
        assert all(torch.equal(a, b) for a, b in zip(r.arena.slots, sl))
        assert float(r.arena.g32.abs().sum()) == 0.0
    with ThreadPoolExecutor(n) as ex:  # and training continues
        [f.result() for f in [ex.submit(grp.replicas[i].step, x[3 * i:3 * i + 3], y[3 * i:3 * i + 3],
                                        fd.logitcrossentropy) for i in range(n)]]
    assert grp.ensure_synced()


def test_verify_exchange_catches_a_broken_exchange():
    """Fault injection: a communicator that forgets one replica's contribution on part of the arena, and one that skips
    the update of a slice, are both reported."""
    from fluxdist.parallel.debug import ExchangeError
    from fluxdist.parallel.fused import ThreadComm

    torch.manual_seed(0)
    base = fd.tiny_convnet(10)
    n = 2
    grp = FusedGroup.for_threads([copy.deepcopy(base) for _ in range(n)], fd.fake_devices(n), fd.Momentum(),
                                 dtype=torch.float32)
    assert grp.verify_exchange()["ok"]
    good = ThreadComm.allreduce_update

    def drops_a_rank(self, rep):  # rank 1's gradient of the first 100 elements never arrives
        keep = self.replicas[1].arena.g32[:100].clone()
        if rep.rank == 0:
            self.replicas[1].arena.g32[:100] = self.replicas[0].arena.g32[:100]
        good(self, rep)
        if rep.rank == 0:
            self.replicas[1].arena.g32[:100] = keep

    def skips_a_slice(self, rep):
        good(self, rep)
        if rep.rank == 1:
            rep.arena.w32[200:300] = float("nan")  # as if nobody had broadcast this slice to replica 1

    for broken, needle in ((skips_a_slice, "never written"), (drops_a_rank, "differs")):
        ThreadComm.allreduce_update = broken
        try:
            with pytest.raises(ExchangeError, match=needle):
                grp.verify_exchange()
            rep = grp.verify_exchange(raise_on_error=False)
            assert not rep["ok"] and rep["errors"]
        finally:
            ThreadComm.allreduce_update = good
    assert grp.verify_exchange()["ok"]


def _gloo_verify_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(10)
    grp = FusedGroup.for_process(fd.tiny_convnet(10), fd.FakeDevice(rank), fd.ADAM(1e-2), dtype=torch.float32, comm="dist")
    rep = grp.verify_exchange()
    q.put((rank, rep["ok"], rep["checked_elems"], rep["max_rel_err"]))
    dist.destroy_process_group()


def test_verify_exchange_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_verify_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(30) for p in ps]
    assert all(ok and n > 0 and err < 1e-4 for _, ok, n, err in res), res
