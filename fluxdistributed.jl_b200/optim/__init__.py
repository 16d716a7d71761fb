"""Functional optimisers (reference L4 / C11, ``src/overloads.jl:1-41``)."""
from .rules import (ADAM, Adam, AdamW, Descent, Momentum, Nesterov, Rule, flat_step, kernel_hyper,
                    state, update)

__all__ = ["ADAM", "Adam", "AdamW", "Descent", "Momentum", "Nesterov", "Rule", "flat_step",
           "kernel_hyper", "state", "update"]
