import torch

import fluxdist as fd
from fluxdist.tree import compare, gradient
from fluxdist.utils import load_checkpoint, save_checkpoint


def test_roundtrip(tmp_path):
    m = fd.tiny_convnet(10)
    opt = fd.ADAM()
    st = fd.state(opt, m)
    g, _ = gradient(lambda mm: mm(torch.rand(2, 3, 32, 32)).sum(), m)
    m, st = opt(m, g, st)
    p = save_checkpoint(str(tmp_path / "a" / "ck.pt"), m, st, step=7, extra={"perm": [1, 2]})
    m2 = fd.tiny_convnet(10)
    ck = load_checkpoint(p, m2)
    assert ck["step"] == 7 and ck["extra"]["perm"] == [1, 2]
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    assert compare(ck["opt_state"], st)
    # resume: continuing from the checkpoint equals continuing in memory
    g2, _ = gradient(lambda mm: mm(torch.ones(2, 3, 32, 32)).sum(), m)
    g3, _ = gradient(lambda mm: mm(torch.ones(2, 3, 32, 32)).sum(), m2)
    m, _ = opt(m, g2, st)
    m2, _ = opt(m2, g3, ck["opt_state"])
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.allclose(a, b, atol=1e-7)
