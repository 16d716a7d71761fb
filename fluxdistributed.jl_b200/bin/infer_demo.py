"""Inference demo (reference C26: ``bin/pluto.jl`` — load a trained model, preprocess a captured
image, print the top-3 labels).  The Pluto/JuliaHub/webcam plumbing has no equivalent here; the
model-side pipeline (``pluto.jl:107-130,379-382``) is reproduced:

    python -m fluxdist.bin.infer_demo --checkpoint ck.pt --image cat.jpg [--labels LOC_synset_mapping.txt]
"""
from __future__ import annotations

import argparse
import sys


def rgba_bytes_to_rgb(buf: bytes, width: int, height: int):
    """RGBA byte stream (what the notebook's canvas hands over, ``pluto.jl:338-369``) -> HxWx3 uint8."""
    import numpy as np

    return np.frombuffer(buf, dtype=np.uint8).reshape(height, width, 4)[:, :, :3].copy()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--image", required=True)
    ap.add_argument("--model", default="resnet34")
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--labels", default=None)
    ap.add_argument("--k", type=int, default=3)
    args = ap.parse_args(argv)
    import torch

    import fluxdist as fd
    from fluxdist import data as D
    from fluxdist.bin.driver import build_model
    from fluxdist.utils import load_checkpoint, maxk

    model = build_model(args.model, args.classes).eval()
    load_checkpoint(args.checkpoint, model)
    probs = D.forward(model, D.load_img(args.image))  # preprocess -> forward -> softmax (pluto.jl:130,379-381)
    names = {}
    if args.labels:
        with open(args.labels) as f:
            for i, line in enumerate(f):
                names[i] = line.strip().partition(" ")[2]
    for idx, p in maxk(probs, args.k):
        print(f"{p:8.4f}  {names.get(idx, f'class {idx}')}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
