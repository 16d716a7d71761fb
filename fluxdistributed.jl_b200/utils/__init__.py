"""Metrics, logging, tracing, checkpointing (reference L2 and SURVEY §5)."""
from .checkpoint import cycle_checkpoint_name, load_checkpoint, save_checkpoint
from .loggers import (ConsoleLogger, JSONLLogger, Logger, MemoryLogger, NullLogger, WandbLogger,
                      get_logger, set_logger, with_logger)
from .metrics import kacc, log_loss_and_acc, maxk, showpreds, topk, topkaccuracy
from .tracing import ClockSampler, DeviceTimer, PhaseTimes, nvtx_range
