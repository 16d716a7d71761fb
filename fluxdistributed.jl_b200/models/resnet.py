"""Model zoo (reference E1: Metalhead ``ResNet34()`` / ``ResNet()``; tiny test net
``test/single_device.jl:119``).

ResNet-18/34 use BasicBlock, ResNet-50/101/152 Bottleneck (v1.5: stride on the 3x3).
Parameter counts match SURVEY App. A: ResNet-34 21,797,672; ResNet-50 25,557,032.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from ..nn.layers import Conv2dN, ConvBN, Head, LinearN, MaxPool


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin: int, planes: int, stride: int = 1):
        super().__init__()
        self.conv1 = ConvBN(cin, planes, 3, stride, relu=True)
        self.conv2 = ConvBN(planes, planes, 3, 1, relu=True)  # relu applied after the residual add
        self.downsample = None
        if stride != 1 or cin != planes:
            self.downsample = ConvBN(cin, planes, 1, stride, padding=0, relu=False)

    def forward(self, x):
        if self.downsample is None and self.conv1.stride == 1 and x.is_cuda and x.dtype == torch.bfloat16 \
                and torch.is_grad_enabled() and self.conv1.cin % 64 == 0:
            # identity skip: the residual gradient is summed into conv1's dgrad epilogue (see ops/native.py)
            link, mid = {}, {}
            # `mid`: conv1's output is consumed by conv2 only, so conv2's dgrad epilogue may do the reduction pass of
            # conv1's BN backward (mask + per-channel sums) and hand over the masked gradient
            from ..ops import native as _nv

            if _nv.DZ_CHAIN and getattr(self, "_chain_ok", False):
                # experimental: the previous block's output is consumed by this block only (guaranteed by ResNet, which
                # sets _chain_ok), so conv1's dgrad may also finish the previous block's last BN layer ("tail" link)
                tail = {"tail": True}
                out = self.conv2(self.conv1(x, add_link=link, own_link=mid, src_link=getattr(x, "_fd_tail", None)),
                                 residual=x, res_link=link, src_link=mid, own_link=tail)
                out._fd_tail = tail
                return out
            return self.conv2(self.conv1(x, add_link=link, own_link=mid), residual=x, res_link=link, src_link=mid)
        if self.downsample is not None and self.conv1.stride == 2 and self.downsample.stride == 2 and x.is_cuda \
                and x.dtype == torch.bfloat16 and torch.is_grad_enabled() and self.conv1.cin % 64 == 0 \
                and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
            # projection skip: the 1x1/2 conv parks its (compact) data gradient and the 3x3/2 conv's parity-decomposed
            # dgrad adds it in its epilogue; conv1 runs first so that autograd schedules the projection's backward first
            link, mid = {"expect_compact": True}, {}
            h = self.conv1(x, add_link=link, own_link=mid)
            return self.conv2(h, residual=self.downsample(x, park_link=link), src_link=mid)
        idn = x if self.downsample is None else self.downsample(x)
        return self.conv2(self.conv1(x), residual=idn)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, cin: int, planes: int, stride: int = 1):
        super().__init__()
        cout = planes * 4
        self.conv1 = ConvBN(cin, planes, 1, 1, padding=0, relu=True)
        self.conv2 = ConvBN(planes, planes, 3, stride, relu=True)
        self.conv3 = ConvBN(planes, cout, 1, 1, padding=0, relu=True)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = ConvBN(cin, cout, 1, stride, padding=0, relu=False)

    def forward(self, x):
        native = x.is_cuda and x.dtype == torch.bfloat16 and torch.is_grad_enabled() and self.conv1.cin % 64 == 0 \
            and self.conv1.cout % 64 == 0
        if native:
            # the same cross-layer links as BasicBlock (ops/native.py::_ConvBNAct): `mid1` / `mid2` let the dgrad of the
            # consumer conv apply the producer BN layer's ReLU mask and reduce its backward sums (dz epilogue: conv3's
            # 1x1 dgrad finishes conv2's BN, conv2's 3x3 dgrad — when stride 1 — finishes conv1's BN)
            mid1, mid2 = {}, {}
            if self.downsample is None and self.conv2.stride == 1:
                # identity skip: conv3 parks the residual gradient, conv1's dgrad epilogue sums it into dx; with the
                # chain enabled conv1 also finishes the previous block's last BN layer and conv3 offers its own
                from ..ops import native as _nv

                link = {}
                if _nv.DZ_CHAIN and getattr(self, "_chain_ok", False):
                    tail = {"tail": True}
                    h = self.conv1(x, add_link=link, own_link=mid1, src_link=getattr(x, "_fd_tail", None))
                    out = self.conv3(self.conv2(h, src_link=mid1, own_link=mid2), residual=x, res_link=link,
                                     src_link=mid2, own_link=tail)
                    out._fd_tail = tail
                    return out
                h = self.conv1(x, add_link=link, own_link=mid1)
                return self.conv3(self.conv2(h, src_link=mid1, own_link=mid2), residual=x, res_link=link, src_link=mid2)
            # projection block: the block input feeds conv1 and the projection, autograd sums their data gradients
            h = self.conv2(self.conv1(x, own_link=mid1), src_link=mid1, own_link=mid2)
            return self.conv3(h, residual=self.downsample(x) if self.downsample is not None else x, src_link=mid2)
        idn = x if self.downsample is None else self.downsample(x)
        return self.conv3(self.conv2(self.conv1(x)), residual=idn)


class ResNet(nn.Module):
    def __init__(self, block, layers: Sequence[int], nclasses: int = 1000, in_ch: int = 3):
        super().__init__()
        self.stem = ConvBN(in_ch, 64, 7, 2, padding=3, relu=True)
        self.pool = MaxPool(3, 2, 1)
        cin = 64
        stages: List[nn.Module] = []
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            blocks = []
            for j in range(n):
                blocks.append(block(cin, planes, stride=(2 if (i > 0 and j == 0) else 1)))
                cin = planes * block.expansion
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3, self.layer4 = stages
        for stage in stages:  # inside a stage every block output feeds the next block only (see BasicBlock.forward)
            for blk in stage:
                blk._chain_ok = True
        self.head = Head(cin, nclasses)
        self.nclasses = nclasses

    def features(self, x):
        if (self.pool.k, self.pool.stride, self.pool.padding) == (3, 2, 1):
            x = self.stem(x, pool=True)  # BN + ReLU + max-pool in one pass on the native path (ops.bn_relu_pool)
        else:
            x = self.pool(self.stem(x))
        return self.layer4(self.layer3(self.layer2(self.layer1(x))))

    def forward(self, x):
        return self.head(self.features(x))


def ResNet18(nclasses: int = 1000) -> ResNet:
    return ResNet(BasicBlock, (2, 2, 2, 2), nclasses)


def ResNet34(nclasses: int = 1000) -> ResNet:
    return ResNet(BasicBlock, (3, 4, 6, 3), nclasses)


def ResNet50(nclasses: int = 1000) -> ResNet:
    return ResNet(Bottleneck, (3, 4, 6, 3), nclasses)


def ResNet101(nclasses: int = 1000) -> ResNet:
    return ResNet(Bottleneck, (3, 4, 23, 3), nclasses)


def ResNet152(nclasses: int = 1000) -> ResNet:
    return ResNet(Bottleneck, (3, 8, 36, 3), nclasses)


def resnet_trunk_with_head(nfeat_classes: int = 1000, nout: int = 200) -> nn.Module:
    """The process-mode default model: ``Chain(ResNet().layers[1:end-1], Dense(1000, 200))``
    (``src/sync.jl:215``) — a ResNet-50 whose 1000-way output feeds a 200-way Dense."""
    return nn.Sequential(ResNet50(nfeat_classes), LinearN(nfeat_classes, nout))


class TinyConvNet(nn.Module):
    """``Chain(Conv((7,7),3=>3), flatten, Dense(2028,10))`` on 32x32x3
    (``test/single_device.jl:119``): 26*26*3 = 2028 features."""

    def __init__(self, nclasses: int = 10, in_hw: int = 32):
        super().__init__()
        self.conv = Conv2dN(3, 3, 7)      # nn.Conv2d / nn.Linear subclasses: native kernels on CUDA bf16 inputs
        self.flatten = nn.Flatten()
        self.dense = LinearN(3 * (in_hw - 6) ** 2, nclasses)

    def forward(self, x):
        return self.dense(self.flatten(self.conv(x)))


def tiny_convnet(nclasses: int = 10) -> TinyConvNet:
    return TinyConvNet(nclasses)
