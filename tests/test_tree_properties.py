"""Property tests (hypothesis) of the structural gradient-tree arithmetic (reference C2/C4/C7/C10: ``destruct``,
``_accum``, ``_dodiv``, ``markbuffer!``/``getbuffer!``, ``sync_buffer``) on randomly shaped trees with ``None`` leaves."""
import torch
from hypothesis import given, settings, strategies as st

import fluxdist  # noqa: F401
from fluxdist.parallel.task import sync_buffer
from fluxdist.tree import accum, compare, fmap, getbuffer, leaves, markbuffer, scale


def _shape():
    return st.lists(st.integers(1, 4), min_size=1, max_size=3)


@st.composite
def skeleton(draw, depth=0):
    """A nested dict of shapes / None (the structure two compatible trees share)."""
    n = draw(st.integers(1, 3))
    out = {}
    for i in range(n):
        kind = draw(st.sampled_from(["tensor", "none", "dict"] if depth < 2 else ["tensor", "none"]))
        out[f"k{i}"] = tuple(draw(_shape())) if kind == "tensor" else (None if kind == "none" else draw(skeleton(depth + 1)))
    return out


def realise(sk, seed):
    g = torch.Generator().manual_seed(seed)

    def rec(s):
        if s is None:
            return None
        if isinstance(s, tuple):
            return torch.randn(*s, generator=g)
        return {k: rec(v) for k, v in s.items()}
    return rec(sk)


@settings(max_examples=40, deadline=None)
@given(skeleton(), st.integers(0, 1000))
def test_accum_commutes_and_scale_distributes(sk, seed):
    a, b = realise(sk, seed), realise(sk, seed + 1)
    assert compare(accum(a, b), accum(b, a), 1e-6, 1e-6)
    lhs = scale(accum(a, b), 0.5)
    rhs = accum(scale(a, 0.5), scale(b, 0.5))
    assert compare(lhs, rhs, 1e-6, 1e-6)
    # None leaves stay None, tensor leaves keep their shape
    for (ka, va), (kb, vb) in zip(leaves(lhs), leaves(a)):
        assert ka == kb and ((va is None) == (vb is None))
        if va is not None:
            assert va.shape == vb.shape


@settings(max_examples=30, deadline=None)
@given(skeleton(), st.integers(2, 4), st.integers(0, 1000))
def test_sync_buffer_is_the_leafwise_mean(sk, n, seed):
    trees = [realise(sk, seed + i) for i in range(n)]
    buffer = {i: t for i, t in enumerate(trees)}
    final = sync_buffer(buffer)
    expect = trees[0]
    for t in trees[1:]:
        expect = accum(expect, t)
    expect = scale(expect, 1.0 / n)
    assert compare(final, expect, 1e-6, 1e-6)


@settings(max_examples=30, deadline=None)
@given(skeleton(), st.integers(0, 1000))
def test_markbuffer_getbuffer_roundtrip(sk, seed):
    src = realise(sk, seed)
    buf = fmap(lambda x: None if x is None else torch.zeros_like(x), src)
    markbuffer(buf, src)
    assert compare(buf, src, 0, 0)
    dst = fmap(lambda x: None if x is None else torch.full_like(x, 7.0), src)
    getbuffer(dst, buf)
    assert compare(dst, src, 0, 0)
