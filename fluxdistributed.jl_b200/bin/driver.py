"""Command-line driver (reference C25: ``bin/driver.jl`` + ``bin/main.jl``).

``python -m fluxdist.bin.driver --mode task|process ...`` reproduces the two recipes of the
reference README: task-based training (``prepare_training`` -> ``train``, ``README.md:21-52``)
and process-based training (``syncgrads`` server + ``start`` workers, ``bin/driver.jl:3-41``).
Configuration is by flags that mirror the reference's keyword arguments.
"""
from __future__ import annotations

import argparse
import sys
import threading


def parse(argv=None):
    ap = argparse.ArgumentParser(prog="fluxdist.bin.driver", description=__doc__)
    ap.add_argument("--mode", choices=["task", "process"], default="task")
    ap.add_argument("--model", default="resnet34", choices=["resnet18", "resnet34", "resnet50", "trunk200", "tiny"])
    ap.add_argument("--dataset", default="imagenet_synthetic", help="name in Data.toml, or 'synthetic'")
    ap.add_argument("--classes", type=int, default=1000)
    ap.add_argument("--devices", type=int, default=None, help="number of devices / workers (default: all GPUs, or 2 fake)")
    ap.add_argument("--opt", default=None, choices=[None, "momentum", "adam", "descent", "nesterov", "adamw"])
    ap.add_argument("--eta", type=float, default=None)
    ap.add_argument("--batch", type=int, default=None, help="per-device batch (task: 96, process: 48)")
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--cycle", type=int, default=5, help="steps when --epochs is unset (ddp_tasks.jl:252)")
    ap.add_argument("--cycles", type=int, default=1, help="process mode: outer cycles (sync.jl:100)")
    ap.add_argument("--nsamples", type=int, default=2000, help="process mode: samples drawn per cycle (sync.jl:95)")
    ap.add_argument("--buffersize", type=int, default=5)
    ap.add_argument("--engine", default="auto", choices=["auto", "staged", "fused"])
    ap.add_argument("--transport", default="channel", choices=["channel", "fused"])
    ap.add_argument("--saveweights", action="store_true")
    ap.add_argument("--weights-root", default="weights")
    ap.add_argument("--checkpoint", default=None, help="task mode: checkpoint path (written every --checkpoint-every steps)")
    ap.add_argument("--checkpoint-every", type=int, default=None)
    ap.add_argument("--resume", default=None, help="checkpoint to resume model weights from")
    ap.add_argument("--log", default=None, help="JSONL log file (default: console)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no-spawn", action="store_true", help="process mode: host the workers as threads")
    return ap.parse_args(argv)


def build_model(name: str, nclasses: int):
    import fluxdist as fd

    if name == "tiny":
        return fd.tiny_convnet(nclasses)
    if name == "trunk200":
        return fd.models.resnet_trunk_with_head(1000, nclasses)
    return {"resnet18": fd.ResNet18, "resnet34": fd.ResNet34, "resnet50": fd.ResNet50}[name](nclasses)


def build_opt(name, eta, mode):
    import fluxdist as fd

    name = name or ("momentum" if mode == "task" else "adam")
    table = {"momentum": fd.Momentum, "adam": fd.ADAM, "descent": fd.Descent, "nesterov": fd.Nesterov, "adamw": fd.AdamW}
    opt = table[name]()
    return opt.with_eta(eta) if eta is not None else opt


def main(argv=None) -> int:
    args = parse(argv)
    import torch

    import fluxdist as fd
    from fluxdist import data as D
    from fluxdist.utils import JSONLLogger, load_checkpoint, with_logger

    classes = range(args.classes)
    shape = (3, 32, 32) if args.model == "tiny" else (3, 224, 224)
    if args.dataset == "synthetic":
        key = D.SyntheticKey(100000, nclasses=args.classes, shape=shape)
        val = None
        tree = None
    else:
        tree = D.dataset(args.dataset)
        key = D.train_solutions(tree, "LOC_train_solution.csv", classes)
        val = D.train_solutions(tree, "LOC_val_solution.csv", classes)
    model = build_model(args.model, args.classes)
    if args.resume:
        load_checkpoint(args.resume, model)
    opt = build_opt(args.opt, args.eta, args.mode)
    ndev = args.devices or (torch.cuda.device_count() if torch.cuda.is_available() else 2)
    devs = fd.cuda_devices(ndev) if torch.cuda.is_available() else fd.fake_devices(ndev)
    logger = JSONLLogger(args.log) if args.log else fd.ConsoleLogger()
    with with_logger(logger):
        if args.mode == "task":
            batch = args.batch or 96
            setup, buffer = fd.prepare_training(model, key, devs, opt, batch, epochs=args.epochs, cycle=args.cycle,
                                                classes=classes, buffersize=args.buffersize, dataset=args.dataset,
                                                engine=args.engine, loss=fd.logitcrossentropy)
            mb = None
            if tree is not None:
                mb = lambda v, nsamples, dataset: D.minibatch(tree, v, nsamples=nsamples, class_idx=classes, dataset=dataset)
            out = fd.train(fd.logitcrossentropy, setup, buffer, opt, val=val, minibatch_fn=mb,
                           checkpoint_every=args.checkpoint_every, checkpoint_path=args.checkpoint)
            print(f"trained {len(out)} replicas")
        else:
            batch = args.batch or 48
            models, states = fd.run_distributed(key, model, fd.logitcrossentropy, nworkers=ndev, data_tree=tree,
                                                class_idx=classes, opt=opt, spawn=not args.no_spawn, devices=devs,
                                                verbose=args.verbose, nsamples=args.nsamples, batchsize=batch,
                                                cycles=args.cycles, saveweights=args.saveweights,
                                                weights_root=args.weights_root, model_name=args.model,
                                                transport=args.transport)
            print(f"trained {len(models)} workers")
    return 0


if __name__ == "__main__":
    sys.exit(main())
