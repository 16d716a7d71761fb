"""fluxdist — a B200-native data-parallel training engine with the capabilities and
public API of FluxDistributed.jl (``/root/reference/src/FluxDistributed.jl:11-12``
exports ``minibatch, train_solutions, syncgrads, prepare_training``).

Layers:  drivers (``bin/``) -> engines (``parallel/``) -> gradient
exchange (``parallel/fused.py`` + ``csrc/``) -> optimisers (``optim/``) -> data
(``data/``) -> metrics/logging (``utils/``) -> device shim (``devices.py``).
"""
__version__ = "0.1.0"

from . import data, devices, losses, models, nn, ops, optim, parallel, tree, utils  # noqa: F401
from .data import (DataLoader, SyntheticKey, center_crop, labels, minibatch, preprocess,  # noqa: F401
                   resize_smallest_dimension, synthetic_minibatch, train_solutions)
from .devices import (CudaDevice, FakeDevice, cuda_devices, device_ctx, fake_devices, synchronize,  # noqa: F401
                      to_cpu, to_device, virtual_devices)
from .losses import logitcrossentropy, loss, onecold, onehot  # noqa: F401
from .models import ResNet18, ResNet34, ResNet50, TinyConvNet, tiny_convnet  # noqa: F401
from .optim import ADAM, Adam, AdamW, Descent, Momentum, Nesterov, state, update as optimiser_update  # noqa: F401
from .parallel import (Channel, RemoteChannel, getgrads, make_channels, prepare_training,  # noqa: F401
                       run_distributed, start, sync_buffer, syncgrads, train, train_step, update)
from .tree import (check_nans, compare, destruct, ensure_synced, getbuffer, getfirst, gradient,  # noqa: F401
                   markbuffer)
from .utils import (ConsoleLogger, MemoryLogger, WandbLogger, log_loss_and_acc, showpreds,  # noqa: F401
                    topkaccuracy, with_logger)

__all__ = ["minibatch", "train_solutions", "syncgrads", "prepare_training", "train", "train_step",
           "sync_buffer", "update", "start", "getgrads", "run_distributed", "destruct", "markbuffer",
           "getbuffer", "ensure_synced", "check_nans", "log_loss_and_acc", "topkaccuracy", "showpreds",
           "Momentum", "ADAM", "Descent", "Nesterov", "state", "logitcrossentropy", "loss"]
