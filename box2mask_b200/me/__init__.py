"""Drop-in operator surface: this package is importable as `MinkowskiEngine`
(call box2mask_b200.install_as_minkowski_engine() before importing the reference's models).

It exports exactly the names the reference uses (94 call sites):
SparseTensor, TensorField, MinkowskiConvolution, MinkowskiConvolutionTranspose, MinkowskiBatchNorm,
MinkowskiSyncBatchNorm, MinkowskiReLU, MinkowskiGlobalAvgPooling, MinkowskiGlobalMaxPooling, cat,
utils.batched_coordinates, utils.kaiming_normal_, modules.resnet_block.Bottleneck, plus import-only
stubs. All computation goes through the C-ABI CUDA library; importing this package does NOT create a
CUDA context (utils.batched_coordinates runs inside DataLoader worker processes,
/root/reference/models/dataloader.py:966).
"""
from .sparse_tensor import CoordinateManager, SparseTensor, TensorField, cat  # noqa: F401
from .nn import (  # noqa: F401
    MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiConvolutionTranspose, MinkowskiGlobalAvgPooling,
    MinkowskiGlobalMaxPooling, MinkowskiReLU, MinkowskiSyncBatchNorm, conv_bn_act, prepack_conv_weights,
)
from . import utils  # noqa: F401
from . import modules  # noqa: F401

__version__ = "0.5.4+b2m_b200"


def __getattr__(name):
    # names the reference only mentions inside never-called methods (models/resnet.py:104-137,219-247)
    if name in ("MinkowskiInstanceNorm", "MinkowskiMaxPooling", "MinkowskiDropout", "MinkowskiGELU",
                "MinkowskiLinear", "MinkowskiSinusoidal", "MinkowskiToSparseTensor", "MinkowskiAvgPooling",
                "MinkowskiSumPooling", "MinkowskiPoolingTranspose", "MinkowskiStableInstanceNorm"):
        from .nn import _unsupported_module
        return _unsupported_module(name)
    raise AttributeError("module 'MinkowskiEngine' (box2mask_b200.me) has no attribute %r" % name)
