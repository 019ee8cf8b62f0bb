"""A CPU stand-in for `MinkowskiEngine` backed by the oracle ops.  TEST INFRASTRUCTURE ONLY.

Purpose: let the REFERENCE's own model code (/root/reference/models/detection_net.py, resnet.py, model.py)
run unmodified in the build container, so that golden vectors pin the oracle's restatement of the network
topology and losses (oracle/make_golden.py). MinkowskiEngine 0.5.4 itself cannot be installed offline;
the op semantics this shim implements are those listed in oracle/sparse_ops.py ("parity unpinned").
"""
import math
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import sparse_ops as so
from .selection_net import CoordCache


class SparseTensor:
    def __init__(self, features, coordinates=None, device=None, _cache=None, _stride=1):
        self.F = features
        if _cache is None:
            _cache = CoordCache(coordinates.cpu().numpy())
        self._cache, self._stride = _cache, _stride
        self._C = None

    @property
    def C(self):
        if self._C is None:
            self._C = torch.from_numpy(self._cache.levels[self._stride].copy())
        return self._C

    def _like(self, f, stride=None):
        return SparseTensor(f, _cache=self._cache, _stride=self._stride if stride is None else stride)

    def __iadd__(self, other):
        self.F = self.F + other.F
        return self


TensorField = SparseTensor


class _Conv(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False, dimension=None,
                 expand_coordinates=False):
        super().__init__()
        self.k, self.s = int(kernel_size), int(stride)
        kv = self.k ** 3
        shape = (in_channels, out_channels) if kv == 1 else (kv, in_channels, out_channels)
        n = (out_channels if self.transposed else in_channels) * kv
        self.kernel = nn.Parameter(torch.empty(shape).uniform_(-1 / math.sqrt(n), 1 / math.sqrt(n)))
        self.bias = nn.Parameter(torch.empty(1, out_channels).uniform_(-1 / math.sqrt(n), 1 / math.sqrt(n))) if bias else None

    def forward(self, x):
        cc, ts = x._cache, x._stride
        out_stride = ts
        if self.k == 1:
            nbr = None
        elif self.s == 1:
            nbr = cc.submanifold(ts, self.k)
        elif not self.transposed:
            nbr, out_stride = cc.stride2(ts)[0], 2 * ts
        else:
            nbr, out_stride = cc.stride2(ts // 2)[1], ts // 2
        y = so.sparse_conv(x.F, nbr, self.kernel)
        if self.bias is not None:
            y = y + self.bias
        return x._like(y, out_stride)


class MinkowskiConvolution(_Conv):
    transposed = False


class MinkowskiConvolutionTranspose(_Conv):
    transposed = True


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum)

    def forward(self, x):
        return x._like(self.bn(x.F))


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):
        return module


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._like(torch.relu(x.F))


class _GlobalPool(nn.Module):
    use_max = False

    def forward(self, x):
        ids = x.C[:, 0].long()
        s = int(ids.max()) + 1
        f = so.segment_max(x.F, ids, s) if self.use_max else so.segment_mean(x.F, ids, s)
        coords = torch.zeros((s, 4), dtype=torch.int32)
        coords[:, 0] = torch.arange(s)
        return SparseTensor(f, coords)


class MinkowskiGlobalAvgPooling(_GlobalPool):
    use_max = False


class MinkowskiGlobalMaxPooling(_GlobalPool):
    use_max = True


def cat(a, b):
    return a._like(torch.cat([a.F, b.F], 1))


def _kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    if tensor.dim() == 2:
        fan_in, fan_out = tensor.size(1), tensor.size(0)
    else:
        fan_in, fan_out = tensor.size(1) * tensor.size(0), tensor.size(2) * tensor.size(0)
    std = torch.nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan_in if mode == "fan_in" else fan_out)
    with torch.no_grad():
        return tensor.normal_(0, std)


def _batched_coordinates(coords, dtype=torch.int32):
    out = [torch.cat([torch.full((len(c), 1), b, dtype=dtype), torch.as_tensor(np.asarray(c)).to(dtype)], 1)
           for b, c in enumerate(coords)]
    return torch.cat(out, 0)


def install():
    """Register this shim as `MinkowskiEngine` (+ stub `open3d`) in sys.modules."""
    me = types.ModuleType("MinkowskiEngine")
    for name in ("SparseTensor", "TensorField", "MinkowskiConvolution", "MinkowskiConvolutionTranspose",
                 "MinkowskiBatchNorm", "MinkowskiSyncBatchNorm", "MinkowskiReLU", "MinkowskiGlobalAvgPooling",
                 "MinkowskiGlobalMaxPooling", "cat"):
        setattr(me, name, globals()[name])
    for name in ("MinkowskiInstanceNorm", "MinkowskiMaxPooling", "MinkowskiDropout", "MinkowskiGELU", "MinkowskiLinear",
                 "MinkowskiSinusoidal", "MinkowskiToSparseTensor"):
        setattr(me, name, type(name, (nn.Module,), {}))
    utils = types.ModuleType("MinkowskiEngine.utils")
    utils.kaiming_normal_ = _kaiming_normal_
    utils.batched_coordinates = _batched_coordinates
    me.utils = utils
    modules = types.ModuleType("MinkowskiEngine.modules")
    rb = types.ModuleType("MinkowskiEngine.modules.resnet_block")
    rb.Bottleneck = type("Bottleneck", (nn.Module,), {"expansion": 4})
    modules.resnet_block = rb
    me.modules = modules
    sys.modules.update({"MinkowskiEngine": me, "MinkowskiEngine.utils": utils, "MinkowskiEngine.modules": modules,
                        "MinkowskiEngine.modules.resnet_block": rb})
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    return me
