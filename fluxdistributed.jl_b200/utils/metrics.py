"""Top-k metrics and prediction pretty-printer (reference C17, ``src/utils.jl:20-71``)
and eval logging (C9, ``src/ddp_tasks.jl:128-148``).

Layout note: the reference is column-major ``[classes, batch]``; everything here is
batch-major ``[batch, classes]``.
"""
from __future__ import annotations

import io
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch


def maxk(a: torch.Tensor, k: int) -> List[Tuple[int, float]]:
    """``maxk`` (``utils.jl:20-25``): the k largest entries of a vector as
    ``(index, value)`` pairs, best first."""
    vals, idx = torch.topk(a.float(), k)
    return list(zip(idx.tolist(), vals.tolist()))


def kacc(preds: Sequence[Sequence[Tuple[int, float]]], truth: Iterable[int]) -> float:
    """``kacc`` (``utils.jl:27-37``): fraction of samples whose true index is in
    the predicted top-k list."""
    truth = list(truth)
    corr = sum(1 for p, t in zip(preds, truth) if t in [i for i, _ in p])
    return corr / max(len(truth), 1)


def topkaccuracy(output: torch.Tensor, labels: torch.Tensor, k: int = 1) -> float:
    """``topkaccuracy(output, onehot_labels; k)`` (``utils.jl:39-45``).

    ``labels`` may be one-hot ``[batch, classes]`` or integer ``[batch]``.  Runs on
    whatever device ``output`` lives on (GPU top-k instead of the
    reference's CPU ``partialsortperm``)."""
    if labels.dim() == 2:
        labels = labels.argmax(dim=-1)
    k = min(k, output.shape[-1])
    top = torch.topk(output.float(), k, dim=-1).indices
    hit = (top == labels.to(top.device).view(-1, 1)).any(dim=-1)
    return hit.float().mean().item()


def topk(output: torch.Tensor, k: int = 5) -> List[List[Tuple[int, float]]]:
    """Per-sample top-k lists (``preprocess.jl`` ``topk`` / ``utils.jl`` ``maxk!.``)."""
    return [maxk(row, k) for row in output]


def showpreds(kk, ground: Sequence[int], labels: Dict[int, str], class_idx: Optional[Sequence[int]] = None,
              file=None) -> str:
    """``showpreds`` (``utils.jl:47-71``): per-sample top-k with a check mark on the hit.

    ``kk``: per-sample ``[(index, value)...]``; ``ground``: true indices; ``labels``:
    class id -> description; ``class_idx`` maps model output index -> class id."""
    buf = io.StringIO()
    cmap = (lambda i: class_idx[i]) if class_idx is not None else (lambda i: i)
    for i, (s, g) in enumerate(zip(kk, ground)):
        buf.write(f"Sample {i + 1}: \n")
        t = cmap(g)
        for idx, v in s:
            c = cmap(idx)
            buf.write(f"  {v:.6g}\t{labels.get(c, str(c))}")
            if c == t:
                buf.write("  ✅")
            buf.write("\n")
        buf.write("\n")
    out = buf.getvalue()
    print(out, file=file)
    return out


def log_loss_and_acc(loss, dev_and_model, val, dataset: str = "val", k=(1, 5, 10), logger=None,
                     minibatch_fn=None, nsamples: int = 300):
    """``log_loss_and_acc`` (``ddp_tasks.jl:128-148``).

    * ``val is None`` -> no-op (``:128``)
    * ``val = (x, y)`` -> upload, ONE eval-mode forward (the reference runs two
      train-mode forwards, quirk Q10), loss + top-k accuracy on the device, log
      ``"<dataset>_<dev>" loss=`` and ``"<dataset>_<k>_<dev>" acc=`` records
    * ``val`` a table (anything with ``sample``) -> draw ``nsamples`` rows via
      ``minibatch_fn`` then recurse (``:142-148``)
    """
    from ..devices import device_ctx, as_device
    from .loggers import get_logger

    if val is None:
        return None
    dev, model = dev_and_model
    if not isinstance(val, (tuple, list)):
        if minibatch_fn is None:
            raise ValueError("a table-valued `val` needs minibatch_fn to materialise a batch")
        val = minibatch_fn(val, nsamples=nsamples, dataset=dataset)
    lg = logger or get_logger()
    d = as_device(dev)
    with device_ctx(d), torch.no_grad():
        x = val[0].to(d.torch_device)
        y = val[1].to(d.torch_device)
        was_training = model.training
        model.eval()
        try:
            fw = model(x)
        finally:
            model.train(was_training)
        l = float(loss(fw, y))
        accs = [topkaccuracy(fw, y, k=kk) for kk in k]
    print(f"{dataset}_loss: {l}")
    lg.info(f"{dataset}_{d}", loss=l)
    for kk, a in zip(k, accs):
        lg.info(f"{dataset}_{kk}_{d}", acc=a)
    return l, dict(zip(k, accs))
