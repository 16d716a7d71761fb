"""Mirror of the reference testset "Distributed Gradient accumulation"
(test/single_device.jl:42-62, helper :6-36): grads of sum(m(batch)) equal the
leafwise _accum of the grads of the single-sample slices."""
import copy

import pytest
import torch
from torch import nn

import fluxdist as fd
from fluxdist.tree import _accum, compare, fmap, getfirst, gradient


def check_data_parallel(m, shape, field=None, rtol=1e-4, atol=1e-4):
    batch = torch.rand(3, *shape)
    full, _ = gradient(lambda mm: mm(batch).sum(), m)
    parts = []
    for i in range(3):
        mi = copy.deepcopy(m)
        g, _ = gradient(lambda mm, i=i: mm(batch[i:i + 1]).sum(), mi)
        parts.append(g)
    acc = parts[0]
    for p in parts[1:]:
        acc = fmap(_accum, acc, p) if acc is not None else p
    assert compare(full, acc, rtol, atol)
    if field is not None:
        a, b = getfirst(full, field), getfirst(acc, field)
        assert a is not None and torch.allclose(a, b, rtol=rtol, atol=atol)


def test_conv():
    check_data_parallel(nn.Conv2d(3, 4, 7), (3, 32, 32), "weight")


def test_dense():
    check_data_parallel(nn.Linear(10, 3), (10,), "weight")


def test_chain_dense():
    check_data_parallel(nn.Sequential(nn.Linear(10, 5), nn.Linear(5, 3)), (10,), "weight")


def test_maxpool_all_nothing():
    m = nn.MaxPool2d(3)
    assert fd.destruct(m) is None
    batch = torch.rand(3, 3, 9, 9, requires_grad=False)
    g, _ = gradient(lambda mm: (mm(batch) * torch.ones(1, requires_grad=True)).sum(), m)
    assert g is None


@pytest.mark.parametrize("field", ["weight", "bias"])
def test_batchnorm_eval(field):
    m = nn.BatchNorm2d(3).eval()
    m.running_mean.uniform_(-1, 1)
    m.running_var.uniform_(0.5, 2)
    check_data_parallel(m, (3, 8, 8), field)


@pytest.mark.parametrize("hw", [32, 64])
def test_conv_bn_eval(hw):
    m = nn.Sequential(nn.Conv2d(3, 4, 7), nn.BatchNorm2d(4)).eval()
    check_data_parallel(m, (3, hw, hw), "weight")


def test_convbn_block_eval():
    m = fd.nn.ConvBN(3, 8, 3, 1).eval()
    check_data_parallel(m, (3, 16, 16), "weight")


def test_resnet18_eval():
    # testmode!(ResNet34()) in the reference; ResNet-18 at 64x64 keeps CPU time low
    m = fd.ResNet18(nclasses=10).eval()
    check_data_parallel(m, (3, 64, 64), "weight", rtol=1e-3, atol=1e-3)
