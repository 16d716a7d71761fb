"""Loss functions (reference C3: ``src/ddp_tasks.jl:28``; README ``logitcrossentropy``).

``loss(x, y) = -sum(y .* logsoftmax(x)) / batch`` — mean softmax cross-entropy
against one-hot (or soft) targets.  Integer class labels are accepted too; on
CUDA bf16/fp32 logits with integer labels the fused sm_100a kernel
(``csrc/softmax_ce.cu``: one pass producing loss and dlogits) is used.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def logitcrossentropy(logits: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Mean softmax-CE.  ``logits`` is ``[batch, classes]``; ``y`` is either a
    one-hot / soft ``[batch, classes]`` matrix or an integer ``[batch]`` vector."""
    if y.dtype in (torch.int64, torch.int32, torch.int16, torch.uint8) and y.dim() == 1:
        if logits.is_cuda:
            from .ops import softmax_ce

            return softmax_ce(logits, y)
        return F.cross_entropy(logits.float(), y.long())
    lsm = F.log_softmax(logits.float(), dim=-1)
    return -(y.to(lsm.dtype) * lsm).sum() / logits.shape[0]


loss = logitcrossentropy  # the reference's default `loss`


def onehot(labels: torch.Tensor, num_classes: int, dtype=torch.float32) -> torch.Tensor:
    """``Flux.onehotbatch`` (batch-major)."""
    return F.one_hot(labels.long(), num_classes).to(dtype)


def onecold(y: torch.Tensor) -> torch.Tensor:
    """``Flux.onecold``."""
    return y.argmax(dim=-1)
