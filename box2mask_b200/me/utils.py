"""ME.utils subset used by the reference. CPU-only helpers; no CUDA context is created here."""
import math

import torch

from ..synthetic import batched_coordinates  # noqa: F401  (models/dataloader.py:966)


def _fans(tensor):
    """MinkowskiEngine's fan computation: dim 0 is the kernel volume for 3-D kernels; 2-D kernels follow the
    Linear convention (fan_in = size(1), fan_out = size(0))."""
    if tensor.dim() < 2:
        raise ValueError("Fan in and fan out can not be computed for tensor with fewer than 2 dimensions")
    if tensor.dim() == 2:
        return tensor.size(1), tensor.size(0)
    rf = tensor.size(0)
    return tensor.size(1) * rf, tensor.size(2) * rf


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """ME.utils.kaiming_normal_ as called at /root/reference/models/resnet.py:142."""
    fan_in, fan_out = _fans(tensor)
    fan = fan_in if mode == "fan_in" else fan_out
    gain = torch.nn.init.calculate_gain(nonlinearity, a)
    std = gain / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)
