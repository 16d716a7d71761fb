#!/usr/bin/env python
"""Process-based data-parallel training: `bin/driver.jl` of FluxDistributed.jl (`syncgrads` server + `start` workers)
with this package.

    python examples/process_mode.py --tiny --workers 2 --no-spawn     # anywhere; workers hosted as threads
    python examples/process_mode.py --workers 8 --transport fused     # 8 B200s: one process per GPU, gradients stay on
                                                                      # the GPUs (fused NVLink kernel), channels carry
                                                                      # one token per step

`transport="channel"` moves CPU-serialised gradient trees through capacity-1 channels exactly like the reference
(`src/sync.jl:36-88,90-170`); the server averages them and every worker applies the optimiser locally.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--cycles", type=int, default=1)
    ap.add_argument("--nsamples", type=int, default=64)
    ap.add_argument("--batch", type=int, default=None, help="per-worker batch (bin/driver.jl: 48)")
    ap.add_argument("--transport", default="channel", choices=["channel", "fused"])
    ap.add_argument("--no-spawn", action="store_true", help="host the workers as threads of this process")
    a = ap.parse_args(argv)

    import torch

    import fluxdist as fd
    from fluxdist import data as D

    classes = range(10 if a.tiny else 1000)
    model = fd.tiny_convnet(len(classes)) if a.tiny else fd.ResNet34(len(classes))
    key = D.SyntheticKey(10_000, nclasses=len(classes), shape=(3, 32, 32) if a.tiny else (3, 224, 224))
    devs = fd.cuda_devices(a.workers) if torch.cuda.is_available() else fd.fake_devices(a.workers)
    # run_distributed = make_channels + a syncgrads server thread + start(...) + collecting the futures
    models, states = fd.run_distributed(key, model, fd.logitcrossentropy, nworkers=a.workers, data_tree=None,
                                        class_idx=classes, opt=fd.ADAM(), spawn=not a.no_spawn, devices=devs,
                                        nsamples=a.nsamples, batchsize=a.batch or (8 if a.tiny else 48),
                                        cycles=a.cycles, transport=a.transport)
    print(f"trained {len(models)} workers")
    return models, states


if __name__ == "__main__":
    main()
