"""End-to-end prepare_training -> train on fake devices (both engines), which the
reference never tests."""
import copy

import pytest
import torch

import fluxdist as fd
from fluxdist.data import SyntheticKey
from fluxdist.utils import MemoryLogger


def _fixed_mb():
    g = torch.Generator().manual_seed(9)
    x = torch.rand(16, 3, 32, 32, generator=g)
    y = fd.onehot(torch.randint(0, 10, (16,), generator=g), 10)

    def mb(shard, n):
        return x[:n], y[:n]

    return mb, (x, y)


@pytest.mark.parametrize("engine", ["staged", "fused"])
def test_train_reduces_loss_and_keeps_replicas_identical(engine):
    torch.manual_seed(0)
    model = fd.tiny_convnet(10)
    key = SyntheticKey(400, nclasses=10, shape=(3, 32, 32))
    mb, (x, y) = _fixed_mb()
    opt = fd.Momentum(1e-3, 0.9)
    lg = MemoryLogger()
    setup, buffer = fd.prepare_training(model, key, fd.fake_devices(2), opt, 8, cycle=20, buffersize=2,
                                        minibatch_fn=mb, engine=engine, dtype=torch.float32)
    l0 = float(fd.logitcrossentropy(model(x[:8]), y[:8]))
    out = fd.train(fd.logitcrossentropy, setup, buffer, opt, val=(x, y), logger=lg, log_every=5, eval_every=10)
    assert [str(d) for d, _ in out] == ["fake:0", "fake:1"]
    m0, m1 = out[0][1], out[1][1]
    for a, b in zip(m0.parameters(), m1.parameters()):
        assert torch.equal(a, b)
    l1 = float(fd.logitcrossentropy(m0(x[:8]), y[:8]))
    assert l1 < l0
    assert len(lg.named("val_fake:0")) == 2 and len(lg.named("train_1_fake:0")) == 2
    # optimiser state is returned in the reference's per-device dict
    assert set(setup.sts) == set(d for d, _ in setup.ds_and_ms)
    assert fd.getfirst(setup.sts[fd.FakeDevice(0)], "weight") is not None


def test_both_engines_agree():
    res = {}
    for engine in ("staged", "fused"):
        torch.manual_seed(0)
        model = fd.tiny_convnet(10)
        key = SyntheticKey(400, nclasses=10, shape=(3, 32, 32))
        mb, _ = _fixed_mb()
        opt = fd.ADAM()
        setup, buffer = fd.prepare_training(model, key, fd.fake_devices(2), opt, 8, cycle=4, minibatch_fn=mb,
                                            engine=engine, dtype=torch.float32)
        res[engine] = fd.train(fd.logitcrossentropy, setup, buffer, opt)[0][1]
    for a, b in zip(res["staged"].parameters(), res["fused"].parameters()):
        assert torch.allclose(a, b, atol=1e-6)


def test_sched_hook_changes_eta():
    model = fd.tiny_convnet(10)
    key = SyntheticKey(100, nclasses=10, shape=(3, 32, 32))
    mb, _ = _fixed_mb()
    lg = MemoryLogger()
    opt = fd.Momentum(0.01)
    setup, buffer = fd.prepare_training(model, key, fd.fake_devices(1), opt, 4, cycle=4, minibatch_fn=mb,
                                        engine="staged")
    fd.train(fd.logitcrossentropy, setup, buffer, opt, sched=lambda o: o.with_eta(o.eta / 5), sched_every=2, logger=lg)
    etas = [r["eta"] for r in lg.named("sched")]
    assert etas == [0.01 / 5, 0.01 / 25]


def test_legacy_argmin_strategy():
    from fluxdist.parallel.legacy import run_distributed_argmin

    model = fd.tiny_convnet(10)
    mb, (x, y) = _fixed_mb()
    m = run_distributed_argmin(fd.logitcrossentropy, model, lambda i: (x, y), (x, y), fd.fake_devices(2), cycles=1,
                               opt=fd.ADAM(1e-3), batchsize=8, logger=MemoryLogger())
    assert isinstance(m, torch.nn.Module)


# --------------------------------------------------------------------------- round 2: resume, collective skip, status
@pytest.mark.parametrize("engine", ["staged", "fused"])
def test_checkpoint_resume_restores_optimiser_state_and_shards(engine, tmp_path):
    """`prepare_training(..., resume=ckpt)`: model weights, optimiser moments + step, shard permutation come back
    (the reference resumes by passing `sts=` back in, src/sync.jl:101,127-129); a resumed run continues exactly like an
    uninterrupted one."""
    mb, _ = _fixed_mb()
    key = SyntheticKey(400, nclasses=10, shape=(3, 32, 32))
    opt = fd.ADAM(1e-3)
    path = str(tmp_path / "ck.pt")

    def run(model, cycle, resume=None, ckpt_every=None):
        setup, buffer = fd.prepare_training(model, key, fd.fake_devices(2), opt, 8, cycle=cycle, minibatch_fn=mb,
                                            engine=engine, dtype=torch.float32, seed=3, resume=resume)
        out = fd.train(fd.logitcrossentropy, setup, buffer, opt, checkpoint_every=ckpt_every, checkpoint_path=path)
        return setup, out[0][1]

    torch.manual_seed(0)
    base = fd.tiny_convnet(10)
    _, full = run(copy.deepcopy(base), 6)                              # uninterrupted: 6 steps
    s1, _ = run(copy.deepcopy(base), 3, ckpt_every=3)                  # interrupted after 3 (checkpoint at step 3)
    ck = torch.load(path, weights_only=False)
    assert ck["step"] == 3 and len(ck["extra"]["shard_indices"]) == 2
    s2, resumed = run(fd.tiny_convnet(10), 3, resume=path)             # fresh model object, resumed from the file
    assert s2.start_step == 3
    assert all((a == b).all() for a, b in zip(s1.shard_indices, s2.shard_indices))
    for a, b in zip(resumed.parameters(), full.parameters()):
        assert torch.allclose(a, b, atol=1e-6), "resumed run diverged from the uninterrupted one"


def test_collective_skip_when_one_replica_fails():
    """A replica that runs out of memory in forward/backward must not leave its peers waiting in the exchange: the
    batch is skipped on EVERY replica, counted in `Num Missed`, and training goes on in sync (src/ddp_tasks.jl:230-237)."""
    torch.manual_seed(0)
    model = fd.tiny_convnet(10)
    key = SyntheticKey(400, nclasses=10, shape=(3, 32, 32))
    mb, (x, y) = _fixed_mb()
    opt = fd.Momentum(1e-3, 0.9)
    calls = {"n": 0}

    def flaky_loss(out, yy):
        calls["n"] += 1
        if calls["n"] == 3:  # second step, first replica to get here
            raise RuntimeError("CUDA out of memory (injected)")
        return fd.logitcrossentropy(out, yy)

    setup, buffer = fd.prepare_training(model, key, fd.fake_devices(2), opt, 8, cycle=5, minibatch_fn=mb,
                                        engine="fused", dtype=torch.float32)
    steps = []
    out = fd.train(flaky_loss, setup, buffer, opt, on_step=lambda j, info: steps.append(j))
    assert len(steps) == 4                      # 5 cycles, one skipped collectively
    for a, b in zip(out[0][1].parameters(), out[1][1].parameters()):
        assert torch.equal(a, b)
    assert setup.engine.ensure_synced()


def test_engine_status_check_raises_on_barrier_timeout():
    """`check_status()` turns the fused kernel's barrier-timeout flag into an exception (round-1: it was write-only)."""
    from fluxdist.parallel.fused import FusedGroup

    class FakeComm:
        name = "fake"

        def __init__(self):
            self.calls = 0

        def status(self, rep):
            self.calls += 1
            return (True, self.calls > 1)

        def close(self):
            pass

    m = fd.tiny_convnet(10)
    grp = FusedGroup.for_threads([m], fd.fake_devices(1), fd.Momentum(), dtype=torch.float32, comm="thread")
    grp.comm = FakeComm()
    assert grp.check_status() == {"nonfinite": True, "barrier_timeout": False}
    with pytest.raises(RuntimeError, match="barrier timed out"):
        grp.check_status()
