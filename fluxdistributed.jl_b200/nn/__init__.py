"""Layer library (reference E2: Flux layers; SURVEY §7.1 ``nn/``)."""
from .layers import (BatchNorm, Chain, Conv, ConvBN, Dense, Flatten, Head, MaxPool, get_backend, set_backend)
