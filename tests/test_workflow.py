"""Mirror of the reference testsets "Workflow" and "Distributed Optimization"
(test/single_device.jl:115-168): mean-reduced distributed grad == full-batch grad,
and a synchronized step leaves every replica equal to the single-device step."""
import copy
from concurrent.futures import ThreadPoolExecutor

import pytest
import torch

import fluxdist as fd
from fluxdist.parallel import sync_buffer, train_step, update
from fluxdist.tree import compare, destruct, gradient


def _world(n, cuda=False):
    if cuda:
        return fd.virtual_devices(n, 0) if torch.cuda.device_count() < n else fd.cuda_devices(n)
    return fd.fake_devices(n)


def _run_workflow(devs, opt):
    torch.manual_seed(1)
    m = fd.tiny_convnet(10)
    x = torch.rand(3 * len(devs), 3, 32, 32)
    y = fd.onehot(torch.randint(0, 10, (3 * len(devs),)), 10)
    loss = fd.logitcrossentropy
    tdev0 = devs[0].torch_device
    reps = [(d, copy.deepcopy(m).to(d.torch_device)) for d in devs]
    buffer = {d: fd.to_device(destruct(m), devs[0]) for d in devs}
    shards = [(x[3 * i:3 * i + 3].to(d.torch_device), y[3 * i:3 * i + 3].to(d.torch_device)) for i, d in enumerate(devs)]
    with ThreadPoolExecutor(len(devs)) as ex:
        gs = [ex.submit(train_step, loss, buffer, d, mm, *s) for (d, mm), s in zip(reps, shards)]
        gs = [g.result() for g in gs]
    final = sync_buffer(buffer)
    mref = copy.deepcopy(m).to(tdev0)
    full, _ = gradient(lambda mm: loss(mm(x.to(tdev0)), y.to(tdev0)), mref)
    assert compare(fd.to_cpu(final), fd.to_cpu(full), 1e-4, 1e-4)
    assert fd.ensure_synced({0: final}, final)
    assert not fd.check_nans(final)

    # Distributed optimisation
    st = fd.state(opt, m)
    sts = {d: fd.to_device(copy.deepcopy(st), d) for d in devs}
    res = [update(opt, dnm, g, final, sts[dnm[0]]) for dnm, g in zip(reps, gs)]
    mref2, stref = opt(mref, full, fd.to_device(copy.deepcopy(st), devs[0]))
    p_ref = [p.detach().cpu() for p in mref2.parameters()]
    for (mm, _st) in res:
        for a, b in zip(mm.parameters(), p_ref):
            assert torch.allclose(a.detach().cpu(), b, rtol=1e-4, atol=1e-4)
    # replicas identical to each other (bitwise)
    for a, b in zip(res[1][0].parameters(), res[2][0].parameters()):
        assert torch.equal(a.detach().cpu(), b.detach().cpu())
    assert compare(fd.to_cpu(res[0][1]), fd.to_cpu(stref), 1e-4, 1e-4)


@pytest.mark.parametrize("opt", [fd.Momentum(), fd.ADAM(), fd.Descent(), fd.Nesterov()])
def test_workflow_fake_devices(opt):
    _run_workflow(_world(3), opt)


@pytest.mark.gpu
@pytest.mark.parametrize("opt", [fd.Momentum(), fd.ADAM()])
def test_workflow_cuda(opt):
    _run_workflow(_world(3, cuda=True), opt)


def test_demo_server_job_picker_and_classify(tmp_path):
    """The notebook demo's pieces (reference bin/pluto.jl): job picker over saved runs, checkpoint selection, and
    classification of a raw RGBA canvas frame."""
    import numpy as np
    import torch

    import fluxdist as fd
    from fluxdist.bin.demo_server import DemoState, create_app
    from fluxdist.utils import save_checkpoint

    root = tmp_path / "weights"
    m = fd.ResNet18(10)
    save_checkpoint(str(root / "0" / "resnet18_cycle_20_x.pt"), m, None, step=20)
    save_checkpoint(str(root / "0" / "resnet18_cycle_40_x.pt"), m, None, step=40)
    st = DemoState("resnet18", 10, str(root), device="cpu")
    jobs = st.jobs()
    assert len(jobs) == 2 and {j["step"] for j in jobs} == {20, 40}
    assert st.select(jobs[0]["path"]).endswith(".pt")
    with pytest.raises(ValueError):
        st.select(str(tmp_path / "elsewhere.pt"))
    frame = np.random.default_rng(0).integers(0, 255, size=(240, 320, 4), dtype=np.uint8).tobytes()
    top = st.classify(frame, 320, 240, k=3)
    assert len(top) == 3 and abs(sum(t["p"] for t in top)) <= 1.0 + 1e-5 and top[0]["p"] >= top[1]["p"]
    app = create_app(st)                                   # routes exist (served with uvicorn by `main`)
    assert {"/", "/jobs", "/select", "/classify"} <= {r.path for r in app.routes}
