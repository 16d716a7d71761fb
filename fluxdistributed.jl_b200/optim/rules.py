"""Functional optimiser rules and the tree-walking glue.

Reference C11 (``src/overloads.jl:1-41``): ``st = state(opt, model)`` builds a
state tree congruent to the model, ``model, st = opt(model, grads, st)`` walks
model/grad/state together, skips ``nothing`` gradients and returns the new
model and state.  The rules follow Optimisers.jl 0.1 as pinned by the reference
(``Manifest.toml:859-865``):

* ``Momentum``: ``v <- rho*v - eta*g ; x <- x + v``   (defaults 0.01, 0.9)
* ``ADAM``:     ``m <- b1*m+(1-b1)g ; v <- b2*v+(1-b2)g^2 ;
                  x <- x - eta * (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)``

Every rule also exposes ``flat_step`` — the same math on flat fp32 buffers —
which is the host-side oracle for the fused all-reduce+update kernel
(``csrc/fused_update.cu``) and the CPU fallback of ``parallel/fused.py``.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch
from torch import nn

from ..tree import Tree, _module_tree

EPS = 1e-8  # Optimisers.jl 0.1 `const ϵ = 1f-8`

# integer ids shared with csrc/fused_update.cu
RULE_DESCENT, RULE_MOMENTUM, RULE_NESTEROV, RULE_ADAM, RULE_ADAMW = 0, 1, 2, 3, 4


class Rule:
    """Base class: a hyper-parameter record plus ``init`` / ``apply``."""

    rule_id: int = -1
    n_slots: int = 0  # number of per-parameter moment buffers

    def init(self, x: torch.Tensor):
        raise NotImplementedError

    def apply(self, state, x: torch.Tensor, dx: torch.Tensor):
        """Return ``(new_state, new_x)``; may update buffers in place."""
        raise NotImplementedError

    # -- functional call convention of the reference: m, st = opt(m, grad, st)
    def __call__(self, model: nn.Module, grads: Tree, st: Tree, inplace: bool = True):
        return update(self, model, grads, st, inplace=inplace)

    def hyper(self) -> Dict[str, float]:
        return {k: v for k, v in self.__dict__.items() if isinstance(v, (int, float))}

    def with_eta(self, eta: float) -> "Rule":
        """Return a copy with a new learning rate (what ``sched(opt)`` would do)."""
        o = copy.copy(self)
        o.eta = eta
        return o


@dataclass
class Descent(Rule):
    eta: float = 0.1
    rule_id = RULE_DESCENT
    n_slots = 0

    def init(self, x):
        return None

    def apply(self, state, x, dx):
        return None, x - self.eta * dx.to(x.dtype)


@dataclass
class Momentum(Rule):
    eta: float = 0.01
    rho: float = 0.9
    rule_id = RULE_MOMENTUM
    n_slots = 1

    def init(self, x):
        return torch.zeros_like(x, dtype=torch.float32)

    def apply(self, state, x, dx):
        v = self.rho * state - self.eta * dx.float()
        return v, (x.float() + v).to(x.dtype)


@dataclass
class Nesterov(Rule):
    eta: float = 0.001
    rho: float = 0.9
    rule_id = RULE_NESTEROV
    n_slots = 1

    def init(self, x):
        return torch.zeros_like(x, dtype=torch.float32)

    def apply(self, state, x, dx):
        g = dx.float()
        v = self.rho * state - self.eta * g
        # Optimisers.jl: dx' = -(rho^2 * v_old) + (1+rho)*eta*g ;  x <- x - dx'
        d = -(self.rho ** 2) * state + (1 + self.rho) * self.eta * g
        return v, (x.float() - d).to(x.dtype)


@dataclass
class ADAM(Rule):
    eta: float = 1e-3
    beta: Tuple[float, float] = (0.9, 0.999)
    eps: float = EPS
    rule_id = RULE_ADAM
    n_slots = 2

    def init(self, x):
        z = torch.zeros_like(x, dtype=torch.float32)
        return (z, z.clone(), (self.beta[0], self.beta[1]))

    def apply(self, state, x, dx):
        mt, vt, bt = state
        b1, b2 = self.beta
        g = dx.float()
        mt = b1 * mt + (1 - b1) * g
        vt = b2 * vt + (1 - b2) * g * g
        d = mt / (1 - bt[0]) / (torch.sqrt(vt / (1 - bt[1])) + self.eps) * self.eta
        return (mt, vt, (bt[0] * b1, bt[1] * b2)), (x.float() - d).to(x.dtype)


@dataclass
class AdamW(ADAM):
    decay: float = 0.0
    rule_id = RULE_ADAMW

    def apply(self, state, x, dx):
        st, xn = ADAM.apply(self, state, x, dx)
        return st, (xn.float() - self.eta * self.decay * x.float()).to(x.dtype)


Adam = ADAM  # spelling used by newer Optimisers.jl


# --------------------------------------------------------------------------- tree glue (C11)
def state(opt: Rule, model: nn.Module) -> Tree:
    """``Optimisers.state(o, x)`` (``overloads.jl:27-34``): recursive ``init``;
    parameter-free leaves give ``None`` (``Optimisers.init(o, x) = nothing``, ``:41``)."""
    return _module_tree(model, lambda p: opt.init(p.detach()) if p.requires_grad else None)


def _walk(opt: Rule, m: nn.Module, g, st, new_state: Dict[str, Any]):
    for name, p in m.named_parameters(recurse=False):
        gi = None if g is None else g.get(name)
        si = None if st is None else st.get(name)
        if gi is None:  # `x̄ === nothing` -> return x, state unchanged (overloads.jl:2-3)
            new_state[name] = si
            continue
        if si is None and opt.n_slots:
            si = opt.init(p.detach())
        s2, x2 = opt.apply(si, p.detach(), gi)
        p.data.copy_(x2)
        new_state[name] = s2
    for name, child in m.named_children():
        gi = None if g is None else g.get(name)
        si = None if st is None else st.get(name)
        sub: Dict[str, Any] = {}
        _walk(opt, child, gi, si, sub)
        new_state[name] = sub if sub else None


def update(opt: Rule, model: nn.Module, grads: Tree, st: Tree, inplace: bool = True):
    """``Optimisers.update(o, x, x̄, state)`` (``overloads.jl:1-12``).

    Returns ``(model, new_state)``.  The reference is out-of-place because Flux
    models are immutable structs; ``nn.Module`` is mutable, so by default the
    returned model is the same object updated in place.  ``inplace=False`` keeps
    the argument untouched and returns an updated deep copy.
    """
    if not inplace:
        model = copy.deepcopy(model)
    new_state: Dict[str, Any] = {}
    with torch.no_grad():
        _walk(opt, model, grads, st, new_state)
    return model, (new_state if new_state else None)


# --------------------------------------------------------------------------- flat oracle for the fused kernel
def flat_step(opt: Rule, w: torch.Tensor, g: torch.Tensor, slots, step: int) -> None:
    """In-place update of flat fp32 master weights ``w`` with mean gradient ``g``.

    ``slots`` is a list of ``opt.n_slots`` fp32 buffers; ``step`` is 1-based and
    supplies the Adam bias-correction powers (beta^step).
    """
    g = g.float()
    if opt.rule_id == RULE_DESCENT:
        w.sub_(opt.eta * g)
    elif opt.rule_id == RULE_MOMENTUM:
        v = slots[0]
        v.mul_(opt.rho).sub_(opt.eta * g)
        w.add_(v)
    elif opt.rule_id == RULE_NESTEROV:
        v = slots[0]
        d = -(opt.rho ** 2) * v + (1 + opt.rho) * opt.eta * g
        v.mul_(opt.rho).sub_(opt.eta * g)
        w.sub_(d)
    elif opt.rule_id in (RULE_ADAM, RULE_ADAMW):
        m, v = slots
        b1, b2 = opt.beta
        m.mul_(b1).add_((1 - b1) * g)
        v.mul_(b2).add_((1 - b2) * g * g)
        c1, c2 = 1 - b1 ** step, 1 - b2 ** step
        d = (m / c1) / (torch.sqrt(v / c2) + opt.eps) * opt.eta
        if opt.rule_id == RULE_ADAMW:
            d = d + opt.eta * opt.decay * w
        w.sub_(d)
    else:
        raise ValueError(f"unknown rule {opt}")


def kernel_hyper(opt: Rule, step: int):
    """Pack hyper-parameters for csrc/fused_update.cu: (rule_id, [h0..h5])."""
    if opt.rule_id == RULE_DESCENT:
        return opt.rule_id, [opt.eta, 0, 0, 0, 0, 0]
    if opt.rule_id in (RULE_MOMENTUM, RULE_NESTEROV):
        return opt.rule_id, [opt.eta, opt.rho, 0, 0, 0, 0]
    b1, b2 = opt.beta
    decay = getattr(opt, "decay", 0.0)
    return opt.rule_id, [opt.eta, b1, b2, opt.eps, 1.0 / (1 - b1 ** step), 1.0 / (1 - b2 ** step), decay][:7]
