import copy

import torch
from torch import nn

import fluxdist as fd
from fluxdist.optim import flat_step
from fluxdist.tree import gradient


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(6, 5), nn.ReLU(), nn.Linear(5, 3))


def test_state_structure_and_none_leaves():
    m = nn.Sequential(nn.Conv2d(3, 3, 3), nn.MaxPool2d(2), nn.Flatten(), nn.Linear(3, 2))
    st = fd.state(fd.Momentum(), m)
    assert st["1"] is None and st["2"] is None
    assert st["0"]["weight"].shape == m[0].weight.shape


def test_momentum_matches_closed_form():
    m = _model()
    x = torch.rand(4, 6)
    opt = fd.Momentum(0.01, 0.9)
    st = fd.state(opt, m)
    w0 = m[0].weight.detach().clone()
    v = torch.zeros_like(w0)
    for _ in range(3):
        g, _ = gradient(lambda mm: mm(x).sum(), m)
        gw = g["0"]["weight"].clone()
        wprev = m[0].weight.detach().clone()
        m, st = opt(m, g, st)
        v = 0.9 * v - 0.01 * gw
        assert torch.allclose(m[0].weight, wprev + v, atol=1e-6)


def test_adam_matches_torch_adam_modulo_eps():
    m = _model()
    m2 = copy.deepcopy(m)
    x = torch.rand(4, 6)
    opt = fd.ADAM(1e-3, (0.9, 0.999))
    st = fd.state(opt, m)
    topt = torch.optim.Adam(m2.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(5):
        g, _ = gradient(lambda mm: (mm(x) ** 2).sum(), m)
        m, st = opt(m, g, st)
        topt.zero_grad()
        (m2(x) ** 2).sum().backward()
        topt.step()
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=2e-5)


def test_nothing_grad_leaves_param_untouched():
    m = _model()
    opt = fd.Momentum()
    st = fd.state(opt, m)
    g, _ = gradient(lambda mm: mm(torch.rand(2, 6)).sum(), m)
    g["2"] = None
    before = m[2].weight.detach().clone()
    m, st = opt(m, g, st)
    assert torch.equal(before, m[2].weight)


def test_out_of_place_update():
    m = _model()
    opt = fd.Descent(0.1)
    g, _ = gradient(lambda mm: mm(torch.rand(2, 6)).sum(), m)
    before = [p.detach().clone() for p in m.parameters()]
    m2, _ = opt(m, g, fd.state(opt, m), inplace=False)
    assert all(torch.equal(a, b) for a, b in zip(before, m.parameters()))
    assert not torch.equal(m2[0].weight, m[0].weight)


def test_flat_step_equals_tree_update():
    for opt in (fd.Momentum(), fd.ADAM(), fd.Descent(), fd.Nesterov(), fd.AdamW(decay=0.01)):
        m = _model()
        st = fd.state(opt, m)
        w = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).clone()
        slots = [torch.zeros_like(w) for _ in range(opt.n_slots)]
        for step in range(1, 4):
            g, _ = gradient(lambda mm: (mm(torch.ones(2, 6)) ** 2).sum(), m)
            gf = torch.cat([g[k][n].reshape(-1) for k in ("0", "2") for n in ("weight", "bias")])
            m, st = opt(m, g, st)
            flat_step(opt, w, gf, slots, step)
            ref = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
            assert torch.allclose(w, ref, atol=1e-6), type(opt).__name__


def _run_rule(opt, steps=4):
    torch.manual_seed(0)
    x = torch.randn(7, 5)
    gs = [torch.randn(7, 5) for _ in range(steps)]
    st = opt.init(x)
    for g in gs:
        st, x = opt.apply(st, x, g)
    return x, gs


def test_descent_matches_torch_sgd():
    x, gs = _run_rule(fd.Descent(0.05))
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(7, 5))
    o = torch.optim.SGD([p], lr=0.05)
    for g in gs:
        p.grad = g.clone()
        o.step()
    assert torch.allclose(x, p.detach(), atol=1e-6)


def test_momentum_matches_torch_sgd_momentum_reparametrised():
    """Optimisers' Momentum (v <- rho v - eta g; x <- x + v) equals torch's (b <- rho b + g; x <- x - lr b) with
    lr = eta (constant learning rate): v = -eta * b."""
    x, gs = _run_rule(fd.Momentum(0.01, 0.9))
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(7, 5))
    o = torch.optim.SGD([p], lr=0.01, momentum=0.9)
    for g in gs:
        p.grad = g.clone()
        o.step()
    assert torch.allclose(x, p.detach(), atol=1e-6)


def test_nesterov_matches_torch_sgd_nesterov():
    x, gs = _run_rule(fd.Nesterov(0.01, 0.9))
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(7, 5))
    o = torch.optim.SGD([p], lr=0.01, momentum=0.9, nesterov=True)
    for g in gs:
        p.grad = g.clone()
        o.step()
    assert torch.allclose(x, p.detach(), atol=1e-6)


def test_adamw_decoupled_decay_matches_torch_adamw_modulo_eps():
    """Same convention difference as ADAM (eps added to sqrt(v_hat) here and in torch; Optimisers 0.1 semantics) — with
    eps tiny the trajectories agree; the decay term is eta*decay*x on the PRE-update weights in both."""
    opt = fd.AdamW(eta=1e-2, beta=(0.9, 0.99), eps=1e-12, decay=0.1)
    x, gs = _run_rule(opt, steps=5)
    torch.manual_seed(0)
    p = torch.nn.Parameter(torch.randn(7, 5))
    o = torch.optim.AdamW([p], lr=1e-2, betas=(0.9, 0.99), eps=1e-12, weight_decay=0.1)
    for g in gs:
        p.grad = g.clone()
        o.step()
    assert torch.allclose(x, p.detach(), atol=2e-5)
