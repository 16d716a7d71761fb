"""Structural gradient trees and their leaf arithmetic.

Reference components C2, C4, C6, C7, C8, C10 (``src/ddp_tasks.jl:4-26, 59-78,
86-126``; ``src/overloads.jl:43-59``).

Zygote hands the reference a NamedTuple mirroring the model's fields with
``nothing`` on non-differentiable leaves; every tree walk there special-cases
``nothing``.  The PyTorch equivalent used throughout this package is a *nested
dict keyed by child-module / parameter name* whose leaves are tensors or
``None`` (frozen parameters, parameter-free modules such as pooling layers).
All walks below are None-aware in the same way.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple

import torch
from torch import nn

Tree = Any  # nested dict / list / tuple with Tensor | None | scalar leaves


# --------------------------------------------------------------------------- walks
def is_leaf(x) -> bool:
    return not isinstance(x, (dict, list, tuple))


def fmap(f: Callable, *trees, none_is_leaf: bool = True) -> Tree:
    """``Functors.fmap`` over congruent trees; ``f`` sees the leaves (incl. None)."""
    t0 = trees[0]
    if isinstance(t0, dict):
        return {k: fmap(f, *[_child(t, k) for t in trees], none_is_leaf=none_is_leaf) for k in t0}
    if isinstance(t0, (list, tuple)):
        out = [fmap(f, *[_child(t, i) for t in trees], none_is_leaf=none_is_leaf) for i in range(len(t0))]
        return type(t0)(out) if not hasattr(t0, "_fields") else type(t0)(*out)
    return f(*trees)


def _child(t, k):
    if t is None:
        return None
    return t[k]


def leaves(tree: Tree, prefix: str = "") -> Iterator[Tuple[str, Any]]:
    """Depth-first ``(dotted_name, leaf)`` pairs (None leaves included)."""
    if isinstance(tree, dict):
        for k, v in tree.items():
            yield from leaves(v, f"{prefix}.{k}" if prefix else str(k))
    elif isinstance(tree, (list, tuple)):
        for i, v in enumerate(tree):
            yield from leaves(v, f"{prefix}.{i}" if prefix else str(i))
    else:
        yield prefix, tree


def tensors(tree: Tree) -> List[torch.Tensor]:
    return [v for _, v in leaves(tree) if isinstance(v, torch.Tensor)]


# --------------------------------------------------------------------------- model <-> tree
def _module_tree(m: nn.Module, leaf_fn: Callable[[nn.Parameter], Any]) -> Tree:
    node: Dict[str, Any] = {}
    for name, p in m.named_parameters(recurse=False):
        node[name] = leaf_fn(p)
    for name, child in m.named_children():
        node[name] = _module_tree(child, leaf_fn)
    if not node:  # parameter-free layer (pooling, activation, flatten): Zygote's `nothing`
        return None
    return node


def destruct(model: nn.Module) -> Tree:
    """Zero gradient skeleton shaped like the model (``ddp_tasks.jl:22-26``).

    Trainable parameters become zero tensors (same device/dtype), frozen
    parameters and parameter-free layers become ``None`` — the ``_zero`` rules of
    ``ddp_tasks.jl:4-9``.
    """
    return _module_tree(model, lambda p: torch.zeros_like(p) if p.requires_grad else None)


def param_tree(model: nn.Module) -> Tree:
    """The live parameters in the same structure (leaves alias the model)."""
    return _module_tree(model, lambda p: p)


def grad_tree(model: nn.Module, detach: bool = True) -> Tree:
    """Collect ``.grad`` into the structural tree (what ``Zygote.gradient`` returns)."""

    def leaf(p):
        if p.grad is None:
            return None
        return p.grad.detach() if detach else p.grad

    return _module_tree(model, leaf)


def gradient(f: Callable[[nn.Module], torch.Tensor], model: nn.Module) -> Tuple[Tree, torch.Tensor]:
    """``gs, = gradient(m -> loss(m(x), y), m)``; also returns the loss value."""
    for p in model.parameters():
        p.grad = None
    out = f(model)
    out.backward()
    gs = grad_tree(model)
    for p in model.parameters():  # the tree owns the buffers now (functional style)
        p.grad = None
    return gs, out.detach()


# --------------------------------------------------------------------------- leaf arithmetic (C10)
def _accum(x, y):
    """``_accum`` (``overloads.jl:43-46``) with the None rules of ``sync_buffer``."""
    if x is None:
        return y
    if y is None:
        return x
    if isinstance(x, torch.Tensor) and x.device != y.device:
        y = y.to(x.device)
    return x + y


def _dodiv(x, n):
    """``_dodiv`` (``overloads.jl:48-54``)."""
    if x is None:
        return None
    return x / n


def _copyto(dst, src):
    """``_copyto!`` (``ddp_tasks.jl:59-64``): None on either side is a no-op."""
    if dst is None or src is None:
        return None
    dst.copy_(src, non_blocking=True)
    return dst


def _zero(x):
    if isinstance(x, torch.Tensor):
        return torch.zeros_like(x)
    return None


def accum(a: Tree, b: Tree) -> Tree:
    return fmap(_accum, a, b) if a is not None else b


def scale(a: Tree, s: float) -> Tree:
    return fmap(lambda x: None if x is None else x * s, a)


# --------------------------------------------------------------------------- buffers (C4)
def markbuffer(dest: Tree, src: Tree, dev=None) -> None:
    """Copy every gradient leaf into the staging buffer for ``dev`` then sync
    (``markbuffer!``, ``ddp_tasks.jl:65-71``)."""
    from .devices import synchronize

    fmap(lambda d, s: _copyto(d, s), dest, src)
    synchronize(dev)


def getbuffer(dest: Tree, src: Tree, dev=None) -> Tree:
    """Copy reduced leaves into the replica's own gradient arrays
    (``getbuffer!``, ``ddp_tasks.jl:73-78``)."""
    fmap(lambda d, s: _copyto(d, s), dest, src)
    return dest


# --------------------------------------------------------------------------- diagnostics (C6, C8, C10)
def check_nans(tree: Tree) -> bool:
    """Recursive ``any(isnan)`` (``ddp_tasks.jl:86-91``)."""
    for _, v in leaves(tree):
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            if bool(torch.isnan(v).any()):
                return True
        elif v != v:
            return True
    return False


def _isapprox(x, y, rtol: float, atol: float) -> bool:
    if x is None and y is None:
        return True
    if x is None or y is None:
        return False
    return bool(torch.allclose(x.to(y.device, y.dtype), y, rtol=rtol, atol=atol))


def ensure_synced(buffer: Dict[Any, Tree], final: Tree, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
    """Every buffer leaf ≈ the corresponding ``final`` leaf (``ddp_tasks.jl:115-126``)."""
    ok = True
    for _dev, g in buffer.items():
        for (_, a), (_, b) in zip(leaves(g), leaves(final)):
            if not _isapprox(a, b, rtol, atol):
                ok = False
    return ok


def show_stats(tree: Tree) -> Tree:
    """``_show_stats`` (``overloads.jl:56-59``): mean/sum/max/min per leaf."""

    def f(x):
        if x is None:
            return None
        xf = x.float()
        st = (xf.mean().item(), xf.sum().item(), xf.max().item(), xf.min().item())
        print("(mean, sum, max, min) =", st)
        return st

    return fmap(f, tree)


def compare(a: Tree, b: Tree, rtol: float = 1e-4, atol: float = 1e-4) -> bool:
    """The test harness' recursive comparison (``test/runtests.jl:6-35``)."""
    la, lb = list(leaves(a)), list(leaves(b))
    if len(la) != len(lb):
        return False
    for (na, x), (nb, y) in zip(la, lb):
        if na != nb:
            return False
        if isinstance(x, torch.Tensor) or isinstance(y, torch.Tensor):
            if not _isapprox(x, y, rtol, atol):
                return False
        elif isinstance(x, (int, float)) and isinstance(y, (int, float)):
            if abs(x - y) > atol + rtol * abs(y):
                return False
        elif x != y:
            return False
    return True


def getfirst(tree: Tree, name: str):
    """Dig the first leaf stored under key ``name`` (``test/runtests.jl:37-41``)."""
    if isinstance(tree, dict):
        if name in tree and not isinstance(tree[name], dict):
            return tree[name]
        for v in tree.values():
            r = getfirst(v, name)
            if r is not None:
                return r
    elif isinstance(tree, (list, tuple)):
        for v in tree:
            r = getfirst(v, name)
            if r is not None:
                return r
    return None


# --------------------------------------------------------------------------- flat views
def flatten_like(model: nn.Module, tree: Tree) -> Optional[torch.Tensor]:
    """Concatenate the tree's tensor leaves in ``model.parameters()`` order."""
    ts = tensors(tree)
    if not ts:
        return None
    return torch.cat([t.reshape(-1) for t in ts])
