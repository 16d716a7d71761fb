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
