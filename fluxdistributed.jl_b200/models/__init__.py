"""Model zoo (replaces Metalhead.jl for the reference's workloads)."""
from .resnet import (BasicBlock, Bottleneck, ResNet, ResNet18, ResNet34, ResNet50, ResNet101, ResNet152,
                     TinyConvNet, resnet_trunk_with_head, tiny_convnet)
