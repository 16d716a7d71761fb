#!/usr/bin/env python
"""Task-based (single node) data-parallel training: the README recipe of FluxDistributed.jl (`README.md:21-52`),
`prepare_training -> train`, with this package.

    python examples/task_mode.py                       # all visible GPUs, ResNet-34, synthetic ImageNet-shaped data
    python examples/task_mode.py --tiny --steps 4      # anywhere (CPU "fake devices" when there is no GPU)

On B200s the fused engine is picked automatically: every replica's step is one CUDA-graph replay of hand-written
sm_100a kernels and the gradient exchange + optimiser is one kernel over NVLink.  Without GPUs the same calls run the
reference algorithm (`engine="staged"`) on CPU replicas, which is what the CPU test-suite exercises.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true", help="tiny ConvNet on 32x32 synthetic data (the reference's test model)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=None, help="per-device batch (README: 96)")
    ap.add_argument("--devices", type=int, default=None)
    a = ap.parse_args(argv)

    import torch

    import fluxdist as fd
    from fluxdist import data as D

    classes = range(10 if a.tiny else 1000)
    model = fd.tiny_convnet(len(classes)) if a.tiny else fd.ResNet34(len(classes))          # model = ResNet34()
    key = D.SyntheticKey(10_000, nclasses=len(classes), shape=(3, 32, 32) if a.tiny else (3, 224, 224))
    # with a real ImageNet tree registered in Data.toml:
    #   tree = D.dataset("imagenet_local"); key = D.train_solutions(tree, "LOC_train_solution.csv", classes)
    opt = fd.Momentum()                                                                       # Optimisers.Momentum()
    if torch.cuda.is_available():
        devs = fd.cuda_devices(a.devices or torch.cuda.device_count())                        # CUDA.devices()
    else:
        devs = fd.fake_devices(a.devices or 2)
    batch = a.batch or (8 if a.tiny else 96)
    setup, buffer = fd.prepare_training(model, key, devs, opt, batch, cycle=a.steps, classes=classes,
                                        dataset="synthetic", loss=fd.logitcrossentropy)
    with fd.with_logger(fd.ConsoleLogger()):
        replicas = fd.train(fd.logitcrossentropy, setup, buffer, opt, val=None, sched=lambda o: o)
    print(f"trained {len(replicas)} replicas for {a.steps} steps on {[str(d) for d in devs]}")
    return replicas


if __name__ == "__main__":
    main()
