"""SelectionNet ("Res16UNet34C" widths + three extra 256-channel levels) built on the b2m operator surface.

Same topology, attribute names and state-dict keys as the reference network
(/root/reference/models/detection_net.py:12-230 built from models/resnet.py:46-83,148-181), so checkpoints are
interchangeable; defined from a stage table instead of straight-line code, and the forward pass
(/root/reference/models/detection_net.py:234-364) calls the fused conv -> BN(+residual)(+ReLU) path.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import functional as Fn
from . import me as ME
from .me.nn import conv_bn_act

PLANES = (32, 64, 128, 256, 256, 128, 96, 96)
ADDED_PLANES = (256, 256, 256, 256, 256, 256)
INIT_DIM = 32

# (strided conv, its BatchNorm, residual stage, stage width)
ENCODER = (
    ("conv1p1s2", "bn1", "block1", PLANES[0]),
    ("conv2p2s2", "bn2", "block2", PLANES[1]),
    ("conv3p4s2", "bn3", "block3", PLANES[2]),
    ("conv4p8s2", "bn4", "block4", PLANES[3]),
    ("added_conv1p16s2", "added_bn1", "added_block1", ADDED_PLANES[0]),
    ("added_conv2p32s2", "added_bn2", "added_block2", ADDED_PLANES[1]),
    ("added_conv3p64s2", "added_bn3", "added_block3", ADDED_PLANES[2]),
)
# (transposed conv, its BatchNorm, residual stage, stage width, index of the encoder output that is concatenated;
#  -1 = the stem output)
DECODER = (
    ("added_convtr4p128s2", "added_bntr4", "added_block4", ADDED_PLANES[3], 5),
    ("added_convtr5p64s2", "added_bntr5", "added_block5", ADDED_PLANES[4], 4),
    ("added_convtr6p32s2", "added_bntr6", "added_block6", ADDED_PLANES[5], 3),
    ("convtr4p16s2", "bntr4", "block5", PLANES[4], 2),
    ("convtr5p8s2", "bntr5", "block6", PLANES[5], 1),
    ("convtr6p4s2", "bntr6", "block7", PLANES[6], 0),
    ("convtr7p2s2", "bntr7", "block8", PLANES[7], -1),
)
# head name in cfg.network_heads -> (attribute name, output width or None = number of classes)
HEADS = {
    "mlp_offsets": ("mlp_offsets", 3),
    "mlp_bounds": ("mlp_bounds", 3),
    "mlp_bb_scores": ("mlp_score", 1),
    "mlp_center_scores": ("mlp_center_score", 1),
    "mlp_semantics": ("mlp_semantics", None),
    "mlp_per_vox_semantics": ("mlp_per_vox_semantics", None),
}


def default_config(**overrides):
    """The fields of the reference config (config_loader.py:11-357) that the hot path reads, with the
    values of configs/scannet.txt."""
    cfg = SimpleNamespace(
        layers=2, in_channels=6, load_unused_head=False, do_segment_pooling=True,
        max_pool_segments_detection_net=False, mlp_bounds_relu=False,
        network_heads=["mlp_offsets", "mlp_bounds", "mlp_bb_scores", "mlp_semantics"],
        mlp_offsets="mlp_offsets", mlp_bounds="mlp_bounds", mlp_bb_scores="mlp_bb_scores",
        mlp_center_scores="mlp_center_scores", mlp_semantics="mlp_semantics",
        mlp_per_vox_semantics="mlp_per_vox_semantics",
        min_bb_size=0.04, loss_on_fg_instances=True, bb_supervision=True, use_bb_iou_loss=False,
        loss_weight_bb_offsets=1.0, loss_weight_bb_bounds=0.5, loss_weight_bb_scores=1.0,
        loss_weight_semantics=1.0, loss_weight_center_scores=None, loss_weight_bb_iou=None,
        loss_weight_per_vox_semantics=1.0, mlp_bb_scores_start_epoch=100, mlp_center_scores_start_epoch=0,
        eval_ths=[0.5, 0.05, 0.3, 0.6], multigpu=False, batch_size=8, voxel_size=0.02,
        # not a reference field: run the U-Net trunk on the hand-scheduled executor (box2mask_b200/trunk.py) instead of
        # module by module; numerics are the same (tests/test_gpu_net.py compares the two)
        trunk_executor=True,
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


class BasicBlock(nn.Module):
    """conv3-BN-ReLU-conv3-BN (+ residual, 1x1 conv + BN when the width changes) - ReLU
    (/root/reference/models/resnet.py:46-83)."""
    expansion = 1

    def __init__(self, inplanes, planes, downsample=None, bn_momentum=0.1, dimension=3):
        super().__init__()
        self.conv1 = ME.MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=1, dimension=dimension)
        self.norm1 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.conv2 = ME.MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dimension=dimension)
        self.norm2 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
        self.relu = ME.MinkowskiReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = conv_bn_act(self.conv1, self.norm1, x, relu=True)
        residual = x if self.downsample is None else conv_bn_act(self.downsample[0], self.downsample[1], x, relu=False)
        return conv_bn_act(self.conv2, self.norm2, out, residual=residual, relu=True)


class SelectionNet(nn.Module):
    def __init__(self, cfg, device, semantic_valid_class_ids, is_foreground=None, out_channels=(96, 96, 3), D=3):
        super().__init__()
        self.cfg, self.device, self.D = cfg, device, D
        self.semantic_valid_class_ids = semantic_valid_class_ids
        self.is_foreground = is_foreground
        layers = int(cfg.layers)

        self.inplanes = INIT_DIM
        self.conv0p1s1 = ME.MinkowskiConvolution(cfg.in_channels, self.inplanes, kernel_size=5, dimension=D)
        self.bn0 = ME.MinkowskiBatchNorm(self.inplanes)
        for conv, bn, block, width in ENCODER:
            setattr(self, conv, ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2, dimension=D))
            setattr(self, bn, ME.MinkowskiBatchNorm(self.inplanes))
            setattr(self, block, self._make_layer(width, layers))
        enc_widths = [w for _, _, _, w in ENCODER]
        for conv, bn, block, width, skip in DECODER:
            setattr(self, conv, ME.MinkowskiConvolutionTranspose(self.inplanes, width, kernel_size=2, stride=2, dimension=D))
            setattr(self, bn, ME.MinkowskiBatchNorm(width))
            self.inplanes = width + (INIT_DIM if skip < 0 else enc_widths[skip])
            setattr(self, block, self._make_layer(width, layers))

        if getattr(cfg, "load_unused_head", False):   # parameters kept for checkpoint compatibility, never used
            self.final0 = ME.MinkowskiConvolution(PLANES[7], out_channels[0], kernel_size=1, bias=True, dimension=D)
            self.final0_bn = ME.MinkowskiBatchNorm(out_channels[0])
            self.final1 = ME.MinkowskiConvolution(out_channels[0], out_channels[1], kernel_size=1, bias=True, dimension=D)
            self.final1_bn = ME.MinkowskiBatchNorm(out_channels[1])
            self.final2 = ME.MinkowskiConvolution(out_channels[1], out_channels[2], kernel_size=1, bias=True, dimension=D)
        self.relu = ME.MinkowskiReLU()

        self.network_heads = {}
        self.requires_voxel_outputs = False
        for head in cfg.network_heads:
            attr, width = HEADS[head]
            if width is None:
                width = len(semantic_valid_class_ids)
            module = self._mlp_head(width, out_channels)
            setattr(self, attr, module)
            self.network_heads[head] = module
            if head == "mlp_per_vox_semantics":
                self.requires_voxel_outputs = True
        self.global_avg_pool = ME.MinkowskiGlobalAvgPooling()
        self.global_max_pool = ME.MinkowskiGlobalMaxPooling()
        # The U-Net trunk runs as one hand-scheduled autograd node (box2mask_b200/trunk.py) unless switched off; the
        # module-by-module path below it is the drop-in surface the reference's own SelectionNet class uses.
        self.use_trunk_executor = bool(getattr(cfg, "trunk_executor", False))
        self._init_weights()

    # -- construction helpers ---------------------------------------------------------------------
    def _make_layer(self, planes, blocks):
        downsample = None
        if self.inplanes != planes:
            downsample = nn.Sequential(
                ME.MinkowskiConvolution(self.inplanes, planes, kernel_size=1, stride=1, dimension=self.D),
                ME.MinkowskiBatchNorm(planes))
        stage = [BasicBlock(self.inplanes, planes, downsample=downsample, dimension=self.D)]
        self.inplanes = planes
        stage += [BasicBlock(planes, planes, dimension=self.D) for _ in range(1, blocks)]
        return nn.Sequential(*stage)

    def _mlp_head(self, width, out_channels):
        D = self.D
        return nn.Sequential(
            ME.MinkowskiConvolution(PLANES[7], out_channels[0], kernel_size=1, bias=True, dimension=D),
            ME.MinkowskiReLU(), ME.MinkowskiBatchNorm(out_channels[0]),
            ME.MinkowskiConvolution(out_channels[0], out_channels[1], kernel_size=1, bias=True, dimension=D),
            ME.MinkowskiReLU(), ME.MinkowskiBatchNorm(out_channels[1]),
            ME.MinkowskiConvolution(out_channels[1], width, kernel_size=1, bias=True, dimension=D))

    def _init_weights(self):
        # models/resnet.py:139-146: Kaiming-normal(fan_out, relu) on non-transposed conv kernels, BN to (1, 0)
        for m in self.modules():
            if isinstance(m, ME.MinkowskiConvolution):
                ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
            if isinstance(m, ME.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def coordinate_plan(self):
        """(number of stride-2 levels, [(tensor_stride, kernel_size) of every stride-1 map]) this network asks for."""
        nlev = len(ENCODER)
        return nlev, [(1, self.conv0p1s1.kernel_size)] + [(2 ** l, 3) for l in range(nlev + 1)]

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, x, pooling_ids=None, num_segments=None):
        # all 8 coordinate levels and 16 kernel maps first (the only host syncs of the step), then a sync-free pass
        ME.prepack_conv_weights(self)      # every conv's forward and dgrad weight image, one launch per step
        x.coordinate_manager.wait_ready()  # maps prefetched on a side stream (Model.prefetch_coordinates), if any
        x.coordinate_manager.prepare(*self.coordinate_plan())   # no-op for levels / maps that already exist
        # the executor covers training (forward + backward) and inference; a backward pass through eval-mode BatchNorm
        # (fine-tuning with frozen statistics, the gradient parity tests) goes module by module through autograd
        if self.use_trunk_executor and (self.training or not torch.is_grad_enabled()):
            out = x._like(self.trunk_executor().run(x), 1)
        else:
            out = self.forward_trunk_modules(x)

        outputs = {}
        if self.requires_voxel_outputs:
            outputs["vox_feats"] = out
        if self.cfg.do_segment_pooling:
            assert pooling_ids is not None
            ids = pooling_ids.to(out.F.device, torch.int64).contiguous()
            s = int(num_segments) if num_segments is not None else int(ids.max().item()) + 1
            fn = Fn.SegmentMaxFn if self.cfg.max_pool_segments_detection_net else Fn.SegmentMeanFn
            pooled = fn.apply(out.F.contiguous(), ids, s)
            coords = torch.zeros((s, 4), dtype=torch.int32, device=pooled.device)
            coords[:, 0] = torch.arange(s, dtype=torch.int32, device=pooled.device)
            # one row per superpoint, keyed by its id: unique by construction, so no coordinate validation pass
            out = ME.SparseTensor(pooled, coordinate_manager=ME.CoordinateManager(coords))
        for head in self.cfg.network_heads:
            src = outputs["vox_feats"] if (self.requires_voxel_outputs and "per_vox" in head) else out
            res = self.network_heads[head](src)
            if self.cfg.mlp_bounds_relu and head == self.cfg.mlp_bounds:
                res = self.relu(res)
            outputs[head] = res
        return outputs

    def trunk_executor(self):
        ex = self.__dict__.get("_trunk_executor")
        if ex is None:
            from .trunk import TrunkExecutor
            ex = TrunkExecutor(self)
            self.__dict__["_trunk_executor"] = ex
        return ex

    def forward_trunk_modules(self, x):
        """The trunk module by module (the reference's own forward order, models/detection_net.py:235-337)."""
        stem = conv_bn_act(self.conv0p1s1, self.bn0, x, relu=True)
        out, skips = stem, []
        for conv, bn, block, _ in ENCODER:
            out = conv_bn_act(getattr(self, conv), getattr(self, bn), out, relu=True)
            out = getattr(self, block)(out)
            skips.append(out)
        for conv, bn, block, _, skip in DECODER:
            out = conv_bn_act(getattr(self, conv), getattr(self, bn), out, relu=True)
            out = ME.cat(out, stem if skip < 0 else skips[skip])
            out = getattr(self, block)(out)
        return out

    def get_prediction(self, batch, with_grad=True, to_cpu=False, to_numpy=False, min_size=True):
        """/root/reference/models/detection_net.py:493-521."""
        ctx = torch.enable_grad() if with_grad else torch.no_grad()
        with ctx:
            sin = ME.SparseTensor(batch["vox_features"], batch["vox_coords"], device=self.device)
            pred = self(sin, batch["pooling_ids"].to(self.device))
        pred = {k: (v.F.float().cpu() if to_cpu else v.F.float()) for k, v in pred.items()}
        if min_size and self.cfg.mlp_bounds in pred and self.cfg.min_bb_size is not None:
            pred[self.cfg.mlp_bounds] = torch.clamp(pred[self.cfg.mlp_bounds], min=self.cfg.min_bb_size)
        if to_numpy:
            pred = {k: v.numpy() for k, v in pred.items()}
        return pred
