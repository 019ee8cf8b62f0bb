"""CPU: the reference's OWN SelectionNet class (when /root/reference is mounted) constructs over
box2mask_b200.me installed as `MinkowskiEngine`, with the same state-dict keys/shapes as our network.
No compute is run (there is no CPU path)."""
import os
import sys
import types

import pytest
import torch

import box2mask_b200
from box2mask_b200.selection_net import SelectionNet, default_config

REF = "/root/reference"


def test_me_surface_names():
    me = box2mask_b200.install_as_minkowski_engine()
    import MinkowskiEngine as ME
    assert ME is me
    from MinkowskiEngine.modules.resnet_block import Bottleneck  # noqa: F401
    for name in ("SparseTensor", "TensorField", "MinkowskiConvolution", "MinkowskiConvolutionTranspose",
                 "MinkowskiBatchNorm", "MinkowskiSyncBatchNorm", "MinkowskiReLU", "MinkowskiGlobalAvgPooling",
                 "MinkowskiGlobalMaxPooling", "cat", "MinkowskiInstanceNorm", "MinkowskiLinear"):
        assert hasattr(ME, name), name
    conv = ME.MinkowskiConvolution(6, 32, kernel_size=5, dimension=3)
    assert conv.kernel.shape == (125, 6, 32) and conv.bias is None
    head = ME.MinkowskiConvolution(96, 3, kernel_size=1, bias=True, dimension=3)
    assert head.kernel.shape == (96, 3) and head.bias.shape == (1, 3)
    assert not isinstance(ME.MinkowskiConvolutionTranspose(8, 8, kernel_size=2, stride=2, dimension=3), ME.MinkowskiConvolution)
    bn = ME.MinkowskiBatchNorm(32)
    assert isinstance(bn.bn, torch.nn.BatchNorm1d) and bn.bn.momentum == 0.1
    sync = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(torch.nn.Sequential(bn, ME.MinkowskiReLU()))
    assert isinstance(sync[0], ME.MinkowskiSyncBatchNorm) and sync[0].bn is bn.bn
    bc = ME.utils.batched_coordinates([torch.zeros(2, 3), torch.ones(3, 3)], dtype=torch.int32)
    assert bc.shape == (5, 4) and bc.dtype == torch.int32 and bc[:, 0].tolist() == [0, 0, 1, 1, 1]
    with pytest.raises(Exception):
        ME.SparseTensor(torch.zeros(2, 6), torch.zeros(2, 4, dtype=torch.int32), device="cpu")   # no CPU path


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")
def test_reference_selection_net_constructs_over_our_surface():
    box2mask_b200.install_as_minkowski_engine()
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    sys.path.insert(0, REF)
    try:
        for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[m]
        import models.detection_net as ref_dn
        cfg = default_config()
        ref_net = ref_dn.SelectionNet(cfg, "cpu", list(range(20)), None, out_channels=[96, 96, 6])
        ours = SelectionNet(cfg, "cpu", list(range(20)), out_channels=[96, 96, 6])
        a = {k: tuple(v.shape) for k, v in ref_net.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert list(a.keys()) == list(b.keys()) and a == b
        ours.load_state_dict(ref_net.state_dict())          # checkpoints are interchangeable
        # Kaiming init of non-transposed convs only (models/resnet.py:139-146): std = sqrt(2 / (K * C_out))
        k = ref_net.block8[0].conv1.kernel
        assert abs(float(k.std()) - (2.0 / (27 * 96)) ** 0.5) < 0.1 * (2.0 / (27 * 96)) ** 0.5
    finally:
        sys.path.remove(REF)
