"""Image preprocessing (reference C16, ``src/preprocess.jl:1-80``).

``preprocess``: resize smallest side to 256 (Gaussian low-pass before
down-scaling, ``:30-42``) -> centre-crop 224 (``:45-49``) -> ``(x-mu)/sigma`` with
the ImageNet statistics (``:51-67``).  The reference then multiplies by 255 and
re-standardises per image with ``Flux.normalise`` (quirk Q9); that is available as
``renormalise=True`` but is not the default.

Output layout is CHW float32 by default; ``layout="HWC"`` gives channels-last,
which is what the sm_100a NHWC kernels consume.
"""
from __future__ import annotations

import io
import os
from typing import Optional, Tuple

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _to_array(im) -> np.ndarray:
    if isinstance(im, np.ndarray):
        return im
    if isinstance(im, torch.Tensor):
        return im.numpy()
    return np.asarray(im.convert("RGB"))


def decode_image(data: bytes) -> np.ndarray:
    """``jpeg_decode`` (``imagenet.jl:32``): bytes -> HxWx3 uint8."""
    if data[:6] == b"\x93NUMPY":
        return np.load(io.BytesIO(data))
    try:
        import cv2

        arr = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        if arr is not None:
            return arr[:, :, ::-1]  # BGR -> RGB
    except ImportError:
        pass
    from PIL import Image

    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def load_img(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        return decode_image(f.read())


def resize_smallest_dimension(im, length: int = 256) -> np.ndarray:
    """``resize_smallest_dimension`` (``preprocess.jl:30-42``)."""
    a = _to_array(im)
    h, w = a.shape[:2]
    r = length / min(h, w)
    nh, nw = max(int(round(h * r)), length), max(int(round(w * r)), length)
    try:
        import cv2

        if r < 1:  # Gaussian low-pass, sigma = 0.25/r like ImageTransformations' recipe
            sig = 0.25 / r
            a = cv2.GaussianBlur(a, (0, 0), sigmaX=sig, sigmaY=sig)
        return cv2.resize(a, (nw, nh), interpolation=cv2.INTER_LINEAR)
    except ImportError:
        from PIL import Image

        return np.asarray(Image.fromarray(a).resize((nw, nh), Image.BILINEAR))


def center_crop(im, length: int = 224) -> np.ndarray:
    """``center_crop`` (``preprocess.jl:45-49``)."""
    a = _to_array(im)
    h, w = a.shape[:2]
    t, l = (h - length) // 2, (w - length) // 2
    return a[t:t + length, l:l + length]


def preprocess(im, size: int = 224, resize: int = 256, layout: str = "CHW", renormalise: bool = False,
               dtype=torch.float32) -> torch.Tensor:
    """``preprocess`` (``preprocess.jl:51-67``) -> float tensor, no batch dim."""
    a = center_crop(resize_smallest_dimension(im, resize), size).astype(np.float32) / 255.0
    a = (a - np.asarray(MEAN, np.float32)) / np.asarray(STD, np.float32)
    if renormalise:  # Q9: x255 then per-image standardisation (imagenet.jl:34)
        a = a * 255.0
        a = (a - a.mean()) / (a.std() + 1e-5)
    t = torch.from_numpy(np.ascontiguousarray(a))
    if layout == "CHW":
        t = t.permute(2, 0, 1).contiguous()
    return t.to(dtype)


def preprocess_u8(im, size: int = 224, resize: int = 256) -> torch.Tensor:
    """Resize+crop only, HWC uint8: the payload of the pinned staging ring; the
    normalisation runs on the GPU (``ops.normalize_u8``)."""
    return torch.from_numpy(np.ascontiguousarray(center_crop(resize_smallest_dimension(im, resize), size)))


# --------------------------------------------------------------------------- pretrained weights
def weights_dir() -> str:
    d = os.environ.get("FLUXDIST_WEIGHTS", os.path.join(os.path.expanduser("~"), ".cache", "fluxdist", "weights"))
    os.makedirs(d, exist_ok=True)
    return d


WEIGHTS_URL = os.environ.get("FLUXDIST_WEIGHTS_URL", "")  # base URL of a weight store (http(s):// or file://)


def getweights(name: str, url: Optional[str] = None) -> str:
    """``getweights(name)`` (``preprocess.jl:3-14``): return the cached weight file, downloading it on a miss from
    ``url`` or ``$FLUXDIST_WEIGHTS_URL/<name>`` (atomic write into the cache).  Without a configured URL or without
    network access only a local hit succeeds and the error says where to place the file."""
    p = os.path.join(weights_dir(), name)
    if os.path.exists(p):
        return p
    src = url or (WEIGHTS_URL.rstrip("/") + "/" + name if WEIGHTS_URL else None)
    if src is None:
        raise FileNotFoundError(f"{p} not present and no weight store is configured (set FLUXDIST_WEIGHTS_URL or pass "
                                f"url=); place the file there")
    import tempfile
    import urllib.error
    import urllib.request

    try:
        with urllib.request.urlopen(src, timeout=60) as r:
            data = r.read()
    except (urllib.error.URLError, OSError, ValueError) as e:
        raise FileNotFoundError(f"could not download {src!r} ({e}); place the file at {p}") from e
    fd, tmp = tempfile.mkstemp(dir=weights_dir(), prefix=".w-")
    with os.fdopen(fd, "wb") as f:
        f.write(data)
    os.replace(tmp, p)
    return p


def weights(name: str, model: Optional[torch.nn.Module] = None):
    """``weights(name)`` (``preprocess.jl:16-24``): load (into ``model`` if given)."""
    sd = torch.load(getweights(name), map_location="cpu", weights_only=False)
    if isinstance(sd, dict) and "model" in sd and "format" in sd:
        sd = sd["model"]
    if model is not None:
        model.load_state_dict(sd)
        return model
    return sd


def forward(model: torch.nn.Module, im, **kw) -> torch.Tensor:
    """``forward(model, im)`` (``preprocess.jl``): preprocess one image and run it."""
    with torch.no_grad():
        p = next(model.parameters())
        x = preprocess(im, **kw).unsqueeze(0).to(p.device)
        return torch.softmax(model(x).float(), dim=-1)[0]


# --------------------------------------------------------------------------- dataset helper stubs (preprocess.jl tail)
def make_fname(img_id: str, dataset: str = "val", base: str = "ILSVRC/Data/CLS-LOC") -> str:
    """``make_fname``: path of an image id inside the CLS-LOC layout (same rule as ``imagenet.makepaths``)."""
    from .imagenet import makepaths

    return makepaths(img_id, dataset, base)


def valimgs(tree, csv: str = "LOC_val_solution.csv", classes=None):
    """``valimgs``: the validation table (ImageId, label, class_idx)."""
    from .imagenet import train_solutions

    return train_solutions(tree, csv, classes)


def testimgs(tree, n: int = 100, prefix: str = "ILSVRC2012_test_"):
    """``testimgs``: ids of the (unlabelled) test split."""
    return [f"{prefix}{i + 1:08d}" for i in range(n)]


def ground_truth(key, img_id: str) -> int:
    """``ground_truth``: class index of ``img_id`` in a solutions table."""
    row = key[key["ImageId"] == img_id]
    if len(row) == 0:
        raise KeyError(img_id)
    return int(row["class_idx"].iloc[0])
