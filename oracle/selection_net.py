"""CPU oracle of the whole network (SelectionNet forward + box-vote losses).  TEST INFRASTRUCTURE ONLY.

A functional restatement of /root/reference/models/detection_net.py:234-364 (forward order),
models/resnet.py:70-83 (BasicBlock) and models/model.py:62-210 (losses) over oracle/sparse_ops.py,
driven by a state dict with the reference's key names. PINNED (topology and loss arithmetic):
tests/golden/selection_net_small.npz was produced by running the reference's own
models/detection_net.py + models/resnet.py + models/model.py over an oracle-backed MinkowskiEngine shim
(oracle/me_shim.py, oracle/make_golden.py) and this restatement reproduces it (tests/test_oracle_net.py).
The MinkowskiEngine op semantics underneath stay "parity unpinned" (see oracle/sparse_ops.py).

emulate_bf16=True rounds activations/weights to bfloat16 at the points where the CUDA path stores bf16
(conv inputs and outputs, BatchNorm outputs), so that the comparison isolates real bugs from rounding.
"""
import zlib

import numpy as np
import torch
import torch.nn.functional as F

from . import sparse_ops as so

ENCODER = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"),
           ("conv4p8s2", "bn4", "block4"), ("added_conv1p16s2", "added_bn1", "added_block1"),
           ("added_conv2p32s2", "added_bn2", "added_block2"), ("added_conv3p64s2", "added_bn3", "added_block3")]
DECODER = [("added_convtr4p128s2", "added_bntr4", "added_block4", 5), ("added_convtr5p64s2", "added_bntr5", "added_block5", 4),
           ("added_convtr6p32s2", "added_bntr6", "added_block6", 3), ("convtr4p16s2", "bntr4", "block5", 2),
           ("convtr5p8s2", "bntr5", "block6", 1), ("convtr6p4s2", "bntr6", "block7", 0), ("convtr7p2s2", "bntr7", "block8", -1)]
HEAD_ATTR = {"mlp_offsets": "mlp_offsets", "mlp_bounds": "mlp_bounds", "mlp_bb_scores": "mlp_score",
             "mlp_center_scores": "mlp_center_score", "mlp_semantics": "mlp_semantics",
             "mlp_per_vox_semantics": "mlp_per_vox_semantics"}


def seeded_state_dict(shapes, seed=0):
    """Deterministic parameters from key names (independent of module construction order): kernels
    ~ N(0, 2/fan_out-ish), BN weights ~ U(0.5, 1.5), biases / running means small, running vars ~ U(0.5, 1.5)."""
    sd = {}
    for key, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + seed) & 0x7FFFFFFF)
        shape = tuple(shape)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros(shape, dtype=torch.long)
        elif key.endswith(".kernel"):
            fan = shape[-2] * (shape[0] if len(shape) == 3 else 1)
            sd[key] = torch.randn(shape, generator=g) * (1.5 / np.sqrt(fan))
        elif key.endswith("running_var"):
            sd[key] = torch.rand(shape, generator=g) + 0.5
        elif key.endswith("bn.weight"):
            sd[key] = torch.rand(shape, generator=g) + 0.5
        else:  # conv bias, bn bias, running_mean
            sd[key] = torch.randn(shape, generator=g) * 0.1
    return sd


class CoordCache:
    """Coordinate levels and kernel maps of one batch (numpy, oracle/sparse_ops.py)."""

    def __init__(self, coords):
        self.levels = {1: np.asarray(coords).astype(np.int32)}
        self.sub, self.s2 = {}, {}

    def submanifold(self, ts, k):
        if (ts, k) not in self.sub:
            self.sub[(ts, k)] = so.kernel_map_submanifold(self.levels[ts], ts, k)
        return self.sub[(ts, k)]

    def stride2(self, ts):
        if ts not in self.s2:
            coarse, parent = so.downsample_coords(self.levels[ts], 2 * ts)
            self.levels[2 * ts] = coarse
            self.s2[ts] = so.kernel_map_stride2(self.levels[ts], parent, len(coarse), ts)
        return self.s2[ts]


class OracleNet:
    def __init__(self, sd, cfg, training, emulate_bf16=False, momentum=0.1, eps=1e-5):
        self.sd, self.cfg, self.training, self.emu = sd, cfg, training, emulate_bf16
        self.momentum, self.eps = momentum, eps

    def _r(self, t):
        return so.bf16_round(t) if self.emu else t

    def _rq(self, t):
        """bf16 rounding with a straight-through gradient (so autograd still works in emulation mode)."""
        return t + (so.bf16_round(t.detach()) - t.detach()) if self.emu else t

    def conv(self, name, x, nbr):
        w = self._rq(self.sd[name + ".kernel"])
        y = so.sparse_conv(self._rq(x), nbr, w)
        if name + ".bias" in self.sd:
            y = y + self.sd[name + ".bias"]
        return y

    def bn(self, name, y, residual=None, relu=False, quantize=True):
        p = name + ".bn."
        yq = self._rq(y) if quantize else y      # the CUDA path stores conv outputs in bf16 ...
        if self.training and not self.emu:
            out = F.batch_norm(y, None, None, self.sd[p + "weight"], self.sd[p + "bias"], True, self.momentum, self.eps)
        elif self.training:
            # ... but takes the batch statistics from the fp32 accumulators before rounding
            mean = y.mean(0)
            var = y.var(0, unbiased=False)
            out = (yq - mean) / torch.sqrt(var + self.eps) * self.sd[p + "weight"] + self.sd[p + "bias"]
        else:
            out = F.batch_norm(yq, self.sd[p + "running_mean"], self.sd[p + "running_var"], self.sd[p + "weight"],
                               self.sd[p + "bias"], False, self.momentum, self.eps)
        if residual is not None:
            out = out + residual
        if relu:
            out = torch.relu(out)
        return self._rq(out) if quantize else out

    def block(self, name, x, nbr):
        """BasicBlock, models/resnet.py:70-83."""
        out = self.bn(name + ".norm1", self.conv(name + ".conv1", x, nbr), relu=True)
        residual = x
        if name + ".downsample.0.kernel" in self.sd:
            residual = self.bn(name + ".downsample.1", self.conv(name + ".downsample.0", x, None))
        return self.bn(name + ".norm2", self.conv(name + ".conv2", out, nbr), residual=residual, relu=True)

    def stage(self, name, x, nbr):
        i = 0
        while "%s.%d.conv1.kernel" % (name, i) in self.sd:
            x = self.block("%s.%d" % (name, i), x, nbr)
            i += 1
        return x

    def head(self, name, x):
        """mlp_head, models/detection_net.py:170-194: (1x1+bias, ReLU, BN) x2, 1x1+bias; fp32."""
        def lin(i, t):
            return t @ self.sd["%s.%d.kernel" % (name, i)] + self.sd["%s.%d.bias" % (name, i)]
        x = self.bn("%s.2" % name, torch.relu(lin(0, x)), quantize=False)
        x = self.bn("%s.5" % name, torch.relu(lin(3, x)), quantize=False)
        return lin(6, x)

    def forward(self, coords, feats, pooling_ids):
        cc = CoordCache(coords)
        x = feats.float()
        stem = self.bn("bn0", self.conv("conv0p1s1", x, cc.submanifold(1, 5)), relu=True)
        out, skips, ts = stem, [], 1
        for conv, bn, block in ENCODER:
            nbr_down, _ = cc.stride2(ts)
            out = self.bn(bn, self.conv(conv, out, nbr_down), relu=True)
            ts *= 2
            out = self.stage(block, out, cc.submanifold(ts, 3))
            skips.append(out)
        for conv, bn, block, skip in DECODER:
            ts //= 2
            _, nbr_up = cc.stride2(ts)
            out = self.bn(bn, self.conv(conv, out, nbr_up), relu=True)
            out = torch.cat([out, stem if skip < 0 else skips[skip]], 1)
            out = self.stage(block, out, cc.submanifold(ts, 3))
        outputs = {}
        vox = out
        if self.cfg.do_segment_pooling:
            s = int(pooling_ids.max()) + 1
            out = so.segment_max(out, pooling_ids, s) if self.cfg.max_pool_segments_detection_net \
                else so.segment_mean(out, pooling_ids, s)
        for head in self.cfg.network_heads:
            src = vox if "per_vox" in head else out
            res = self.head(HEAD_ATTR[head], src)
            if self.cfg.mlp_bounds_relu and head == "mlp_bounds":
                res = torch.relu(res)
            outputs[head] = res
        return outputs


def detection_loss(pred, batch, cfg, epoch, semantic_id2idx):
    """optimization_loss of models/model.py:62-210 for the heads of configs/scannet.txt (+ optional IoU loss)."""
    from .nms import iou_aligned
    fg = batch["fg_instances"]
    use_fg = cfg.loss_on_fg_instances or cfg.bb_supervision

    def sel(t):
        return t[fg] if use_fg else t
    total = 0.0
    parts = {}
    if "mlp_offsets" in cfg.network_heads:
        parts["offset_loss"] = torch.mean(torch.sum(torch.abs(sel(pred["mlp_offsets"]) - sel(batch["gt_bb_offsets"])), 1))
        total = total + cfg.loss_weight_bb_offsets * parts["offset_loss"]
    if "mlp_bounds" in cfg.network_heads:
        parts["bounds_loss"] = torch.mean(torch.sum(torch.abs(sel(pred["mlp_bounds"]) - sel(batch["gt_bb_bounds"])), 1))
        total = total + cfg.loss_weight_bb_bounds * parts["bounds_loss"]
    if getattr(cfg, "use_bb_iou_loss", False):
        # optional IoU loss, models/model.py:91-129: 1 - IoU with union clamped from below by 1e-6 (no +eps)
        loc = sel(batch["input_location"])
        pb = torch.clamp(sel(pred["mlp_bounds"]), min=cfg.min_bb_size)
        pc, gc, gb = sel(pred["mlp_offsets"]) + loc, sel(batch["gt_bb_offsets"]) + loc, sel(batch["gt_bb_bounds"])
        pr, gt = torch.cat([pc - pb, pc + pb], 1), torch.cat([gc - gb, gc + gb], 1)
        area1 = (pr[:, 3] - pr[:, 0]) * (pr[:, 4] - pr[:, 1]) * (pr[:, 5] - pr[:, 2])
        area2 = (gt[:, 3] - gt[:, 0]) * (gt[:, 4] - gt[:, 1]) * (gt[:, 5] - gt[:, 2])
        wh = (torch.min(pr[:, 3:], gt[:, 3:]) - torch.max(pr[:, :3], gt[:, :3])).clamp(min=0)
        overlap = wh[:, 0] * wh[:, 1] * wh[:, 2]
        union = torch.max(area1 + area2 - overlap, overlap.new_tensor([1e-6]))
        parts["iou_loss"] = torch.mean(1.0 - overlap / union)
        total = total + cfg.loss_weight_bb_iou * parts["iou_loss"]
    if "mlp_bb_scores" in cfg.network_heads:
        w = cfg.loss_weight_bb_scores if epoch >= cfg.mlp_bb_scores_start_epoch else 0
        loc = sel(batch["input_location"])
        gt_c, gt_b = sel(batch["gt_bb_offsets"]) + loc, sel(batch["gt_bb_bounds"])
        pb = torch.clamp(sel(pred["mlp_bounds"]), min=cfg.min_bb_size)
        pc = sel(pred["mlp_offsets"]) + loc
        ious = iou_aligned(torch.cat([gt_c - gt_b, gt_c + gt_b], 1), torch.cat([pc - pb, pc + pb], 1)).detach()
        parts["bb_score_loss"] = F.binary_cross_entropy_with_logits(sel(pred["mlp_bb_scores"].reshape(-1)), ious)
        parts["bb_target_scores"] = ious.mean()
        total = total + w * parts["bb_score_loss"]
    if "mlp_semantics" in cfg.network_heads:
        gt = semantic_id2idx[batch["gt_semantics"]]
        parts["semantics_loss"] = F.cross_entropy(pred["mlp_semantics"], gt, ignore_index=-100)
        total = total + cfg.loss_weight_semantics * parts["semantics_loss"]
    if "mlp_per_vox_semantics" in cfg.network_heads:       # models/model.py:213-224
        gt = semantic_id2idx[batch["gt_per_vox_semantics"]]
        parts["per_vox_semantics_loss"] = F.cross_entropy(pred["mlp_per_vox_semantics"], gt, ignore_index=-100)
        total = total + cfg.loss_weight_per_vox_semantics * parts["per_vox_semantics_loss"]
    parts["optimization_loss"] = total
    return parts
