"""Parallel engines (reference L6/L5): task-based (``task``), process-based
(``process``), channels, and the fused B200 gradient-exchange engine (``fused``)."""
from .channels import Channel, ChannelClosed, RemoteChannel, StoreChannel, wait_all_ready
from .fused import DistComm, FlatArena, FusedGroup, FusedReplica, ThreadComm
from .process import getgrads, make_channels, run_distributed, start, syncgrads
from .task import Setup, prepare_training, shard_indices, sync_buffer, train, train_step, update
