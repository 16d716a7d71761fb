"""Process-mode tests the reference lacks: 2 workers over the
channel transport, thread-hosted and spawned, CPU."""
import copy
import threading

import pytest
import torch

import fluxdist as fd
from fluxdist.parallel import make_channels, run_distributed, start, syncgrads
from fluxdist.tree import gradient


def _data(n=24):
    g = torch.Generator().manual_seed(3)
    return torch.rand(n, 3, 32, 32, generator=g), fd.onehot(torch.randint(0, 10, (n,), generator=g), 10)


class FixedMB:
    """Deterministic: worker i always gets the same nsamples rows."""

    def __call__(self, tree, key, nsamples):
        return key[0][:nsamples], key[1][:nsamples]


def _expected(model, key, opt, steps, bs):
    m = copy.deepcopy(model)
    st = fd.state(opt, m)
    for s in range(steps):
        x, y = key[0][s * bs:(s + 1) * bs], key[1][s * bs:(s + 1) * bs]
        g, _ = gradient(lambda mm: fd.logitcrossentropy(mm(x), y), m)
        m, st = opt(m, g, st)
    return m


@pytest.mark.parametrize("spawn", [False, True])
def test_two_workers_equal_single(spawn):
    torch.manual_seed(0)
    model = fd.tiny_convnet(10)
    key = _data(8)
    opt = fd.Momentum(1e-3, 0.9)
    models, states = run_distributed(key, model, fd.logitcrossentropy, nworkers=2, opt=opt, spawn=spawn,
                                     nsamples=8, batchsize=4, cycles=1, minibatch_fn=FixedMB(),
                                     devices=fd.fake_devices(2))
    # identical data on both workers -> mean grad == single grad -> same as 1-worker training
    ref = _expected(model, key, opt, 2, 4)
    for m in models:
        for a, b in zip(m.parameters(), ref.parameters()):
            assert torch.allclose(a, b, atol=1e-5)
    for a, b in zip(models[0].parameters(), models[1].parameters()):
        assert torch.equal(a, b)
    assert states[0] is not None


def test_resume_with_states():
    torch.manual_seed(0)
    model = fd.tiny_convnet(10)
    key = _data(8)
    opt = fd.ADAM()
    ms, sts = run_distributed(key, model, nworkers=2, opt=opt, spawn=False, nsamples=4, batchsize=4, cycles=1,
                              minibatch_fn=FixedMB(), devices=fd.fake_devices(2))
    ms2, sts2 = run_distributed(key, ms, nworkers=2, opt=opt, spawn=False, nsamples=4, batchsize=4, cycles=1,
                                minibatch_fn=FixedMB(), devices=fd.fake_devices(2), sts=sts)
    # Adam's beta powers advanced twice -> state really resumed
    bt = fd.getfirst(sts2[0], "weight")[2]
    assert abs(bt[0] - 0.9 ** 3) < 1e-9


def test_checkpoint_written(tmp_path):
    model = fd.tiny_convnet(10)
    key = _data(8)
    run_distributed(key, model, nworkers=2, opt=fd.Momentum(), spawn=False, nsamples=4, batchsize=4, cycles=2,
                    minibatch_fn=FixedMB(), devices=fd.fake_devices(2), saveweights=True, save_every=1,
                    weights_root=str(tmp_path / "weights"), model_name="tiny")
    files = list((tmp_path / "weights" / "0").glob("tiny_cycle_*.pt"))
    assert len(files) >= 1
    ck = fd.utils.load_checkpoint(str(files[0]), fd.tiny_convnet(10))
    assert ck["step"] >= 1
