"""Host wrapper: network + box-vote losses (the reference's models/model.py, restated).

Loss arithmetic follows /root/reference/models/model.py:38-225 term by term. The reference's
diagnostics that force a device->host sync every step (pearsonr models/model.py:170, semIOU / .cpu()
models/model.py:204-210) are computed only when `diagnostics=True`; they do not enter the optimised loss.
"""
import torch

from . import me as ME
from .selection_net import SelectionNet


def to_bbs_min_max_(centers, bounds):
    """/root/reference/utils/util.py:66-70: [centre - bounds, centre + bounds]."""
    return torch.cat([centers - bounds, centers + bounds], dim=1)


def set_ious(boxes_a, boxes_b):
    """Row-wise axis-aligned IoU, /root/reference/models/iou_nms.py:4-22 (same operation order)."""
    side_a = boxes_a[:, 3:] - boxes_a[:, :3]
    side_b = boxes_b[:, 3:] - boxes_b[:, :3]
    lo = torch.maximum(boxes_a[:, :3], boxes_b[:, :3])
    hi = torch.minimum(boxes_a[:, 3:], boxes_b[:, 3:])
    inter = torch.prod(torch.clamp(hi - lo, min=0), dim=1)
    union = torch.prod(side_a, dim=1) + torch.prod(side_b, dim=1) - inter + 0.000001
    return inter / union


class Model:
    def __init__(self, cfg, semantic_valid_class_ids, semantic_id2idx, instance_id2idx=None, is_foreground=None,
                 device="cuda", diagnostics=False):
        self.cfg, self.device, self.diagnostics = cfg, device, diagnostics
        self.semantic_valid_class_ids = semantic_valid_class_ids
        self.semantic_id2idx = semantic_id2idx.to(device) if torch.is_tensor(semantic_id2idx) else semantic_id2idx
        self.instance_id2idx = instance_id2idx
        self.is_foreground = is_foreground
        self.detection_model = SelectionNet(cfg, device, semantic_valid_class_ids, is_foreground,
                                            out_channels=[96, 96, 6]).to(device)
        self.net = self.detection_model   # the un-wrapped network (state-dict owner)
        self.grad_sync = None
        if cfg.multigpu:
            # models/model.py:23-25: gradient averaging over the scene-sharded ranks + SyncBatchNorm.
            # cfg.grad_sync: "overlap" (default) = the deep levels' gradients all-reduced from inside backward on a side
            # stream (grad_sync.TrunkGradSync); "flat" = one all-reduce after backward; "ddp" = torch's bucketed DDP.
            mode = getattr(cfg, "grad_sync", "overlap")
            if mode == "ddp":
                self.net.use_trunk_executor = False      # DDP needs autograd's per-parameter hooks
                self.detection_model = torch.nn.parallel.DistributedDataParallel(
                    self.detection_model, device_ids=[torch.device(device).index], gradient_as_bucket_view=True)
            elif mode == "overlap" and self.net.use_trunk_executor:
                # deep-level gradients all-reduced on a side stream from inside the trunk's backward pass
                from .grad_sync import TrunkGradSync
                self.grad_sync = TrunkGradSync(self.net, free_sms=getattr(cfg, "grad_sync_free_sms", 8))
            else:
                from .grad_sync import FlatGradSync
                self.grad_sync = FlatGradSync(self.net)
            ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(self.net)
        self.bce = torch.nn.BCEWithLogitsLoss()
        self.ce = torch.nn.CrossEntropyLoss(ignore_index=-100)

    # ---------------------------------------------------------------------------------------------
    def compute_loss(self, batch, epoch):
        return self.compute_loss_detection(batch, epoch)[0]

    def prefetch_coordinates(self, batch):
        """Optional: build the coordinate levels and kernel maps of `batch` on a side stream NOW (typically right after
        the previous step was enqueued, so that the integer map construction overlaps its backward pass). The batch
        keeps the result under "_coordinate_manager"; compute_loss_detection / get_prediction pick it up. Purely an
        overlap of work that would otherwise run at the start of the step; results are identical."""
        dev = torch.device(self.device)
        if not hasattr(self, "_side_stream"):
            self._side_stream = torch.cuda.Stream(device=dev)
        side = self._side_stream
        side.wait_stream(torch.cuda.current_stream(dev))     # the coordinates may have been produced on the main stream
        with torch.cuda.stream(side):
            coords = batch["vox_coords"].to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()
        cm = ME.CoordinateManager(coords)
        cm.prepare(*self.net.coordinate_plan(), stream=side)
        batch["_coordinate_manager"] = cm
        return batch

    def stage_batch(self, batch):
        """Optional: copy the tensors of a host batch (pinned memory) to the device on a copy stream NOW - typically
        right after the previous step was enqueued, so that the transfer (59 MB for 8 ScanNet scenes) runs under that
        step's kernels instead of in front of this one's. Returns the device batch; compute_loss_detection /
        get_prediction wait for the copy before they touch it."""
        dev = torch.device(self.device)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._copy_stream):
            out = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        out["_staged"] = done
        return out

    def _wait_staged(self, batch):
        done = batch.pop("_staged", None) if isinstance(batch, dict) else None
        if done is not None:
            main = torch.cuda.current_stream(torch.device(self.device))
            main.wait_event(done)
            for v in batch.values():        # allocated under the copy stream, used on this one
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(main)

    def _input_tensor(self, batch):
        self._wait_staged(batch)
        cm = batch.pop("_coordinate_manager", None) if isinstance(batch, dict) else None
        if cm is not None:
            return ME.SparseTensor(batch["vox_features"].to(self.device), coordinate_manager=cm)
        return ME.SparseTensor(batch["vox_features"], batch["vox_coords"], device=self.device)

    def compute_loss_detection(self, batch, epoch):
        cfg, dev = self.cfg, self.device
        sin = self._input_tensor(batch)
        pred = self.detection_model(sin, batch["pooling_ids"].to(dev), batch.get("num_segments"))
        pred = {k: v.F.float() for k, v in pred.items()}
        losses = {"optimization_loss": 0}
        use_fg = cfg.loss_on_fg_instances or cfg.bb_supervision
        fg = batch["fg_instances"].to(dev)
        # rows of the foreground superpoints as an index list: a boolean-mask gather reads its size back from the device
        # for every tensor it is applied to (ten per step). The list comes with the batch ("fg_index", from the collate
        # function), is computed on the host while the mask still lives there, or costs one read-back.
        fg_index = batch.get("fg_index") if isinstance(batch, dict) else None
        if fg_index is None and use_fg:
            src = batch["fg_instances"]
            fg_index = torch.nonzero(src.reshape(-1)).reshape(-1)
        if fg_index is not None:
            fg_index = fg_index.to(dev)

        def sel(t):
            t = t.to(dev)
            return t.index_select(0, fg_index) if use_fg else t

        offset_loss_per_pred = None
        if cfg.mlp_offsets in cfg.network_heads:
            offset_loss_per_pred = torch.sum(torch.abs(sel(pred[cfg.mlp_offsets]) - sel(batch["gt_bb_offsets"])), dim=1)
            offset_loss = torch.mean(offset_loss_per_pred)
            losses["optimization_loss"] = losses["optimization_loss"] + cfg.loss_weight_bb_offsets * offset_loss
            losses["offset_loss"] = offset_loss.detach()
        if cfg.mlp_bounds in cfg.network_heads:
            bounds_loss = torch.mean(torch.sum(torch.abs(sel(pred[cfg.mlp_bounds]) - sel(batch["gt_bb_bounds"])), dim=1))
            losses["optimization_loss"] = losses["optimization_loss"] + cfg.loss_weight_bb_bounds * bounds_loss
            losses["bounds_loss"] = bounds_loss.detach()
        if cfg.use_bb_iou_loss:
            loc = sel(batch["input_location"])
            pb = torch.clamp(sel(pred[cfg.mlp_bounds]), min=cfg.min_bb_size)
            pr = to_bbs_min_max_(sel(pred[cfg.mlp_offsets]) + loc, pb)
            gt = to_bbs_min_max_(sel(batch["gt_bb_offsets"]) + loc, sel(batch["gt_bb_bounds"]))
            a1 = (pr[:, 3] - pr[:, 0]) * (pr[:, 4] - pr[:, 1]) * (pr[:, 5] - pr[:, 2])
            a2 = (gt[:, 3] - gt[:, 0]) * (gt[:, 4] - gt[:, 1]) * (gt[:, 5] - gt[:, 2])
            wh = (torch.min(pr[:, 3:], gt[:, 3:]) - torch.max(pr[:, :3], gt[:, :3])).clamp(min=0)
            overlap = wh[:, 0] * wh[:, 1] * wh[:, 2]
            union = torch.max(a1 + a2 - overlap, overlap.new_tensor([1e-6]))
            iou_loss = torch.mean(1.0 - overlap / union)
            losses["optimization_loss"] = losses["optimization_loss"] + cfg.loss_weight_bb_iou * iou_loss
            losses["iou_loss"] = iou_loss.detach()
        if cfg.mlp_bb_scores in cfg.network_heads:
            weight = cfg.loss_weight_bb_scores if epoch >= cfg.mlp_bb_scores_start_epoch else 0
            scores = sel(pred[cfg.mlp_bb_scores].reshape(-1))
            loc = sel(batch["input_location"])
            gt_bbs = to_bbs_min_max_(sel(batch["gt_bb_offsets"]) + loc, sel(batch["gt_bb_bounds"]))
            pb = torch.clamp(sel(pred[cfg.mlp_bounds]), min=cfg.min_bb_size)
            pred_bbs = to_bbs_min_max_(sel(pred[cfg.mlp_offsets]) + loc, pb)
            ious = set_ious(gt_bbs, pred_bbs).detach()
            score_loss = self.bce(scores, ious)
            if self.diagnostics:
                from scipy.stats import pearsonr
                losses["bb_scores_correlation"] = pearsonr(ious.cpu().numpy(), scores.detach().cpu().numpy())[0]
            losses["optimization_loss"] = losses["optimization_loss"] + weight * score_loss
            losses["bb_score_loss"] = score_loss.detach()
            losses["bb_target_scores"] = torch.mean(ious)
        if cfg.mlp_center_scores in cfg.network_heads and epoch >= cfg.mlp_center_scores_start_epoch:
            cs = pred[cfg.mlp_center_scores].reshape(-1)
            if cfg.loss_on_fg_instances:
                cs = cs.index_select(0, fg_index) if fg_index is not None else cs[fg]
            cs_loss = torch.mean(torch.abs(cs - offset_loss_per_pred.detach()))
            losses["optimization_loss"] = losses["optimization_loss"] + cfg.loss_weight_center_scores * cs_loss
            losses["center_score_loss"] = cs_loss.detach()
        for head, gt_key, w, tag in ((cfg.mlp_semantics, "gt_semantics", cfg.loss_weight_semantics, "semantics"),
                                     (cfg.mlp_per_vox_semantics, "gt_per_vox_semantics",
                                      cfg.loss_weight_per_vox_semantics, "per_vox_semantics")):
            if head not in cfg.network_heads:
                continue
            logits = pred[head]
            gt = self.semantic_id2idx[batch[gt_key].to(dev)]
            sem_loss = self.ce(logits, gt)
            losses["optimization_loss"] = losses["optimization_loss"] + w * sem_loss
            losses[tag + "_loss"] = sem_loss.detach()
            if self.diagnostics:
                losses[tag + "_acc"] = (torch.sum(torch.argmax(logits, 1) == gt) / len(gt)).detach()
        return losses, pred

    # ---------------------------------------------------------------------------------------------
    def get_prediction(self, batch, with_grad=False, to_cpu=True, min_size=True):
        self._wait_staged(batch)
        return self.net.get_prediction(batch, with_grad=with_grad, to_cpu=to_cpu, min_size=min_size)

    def pred2mask(self, batch, pred, mode="eval"):
        from .decode import detection2mask
        return detection2mask(self.net, batch, pred, self.cfg, mode, True, *self.cfg.eval_ths)

    def parameters(self):
        return self.detection_model.parameters()

    def to(self, device):
        self.detection_model = self.detection_model.to(device)
        return self

    def eval(self):
        self.detection_model.eval()

    def train(self):
        self.detection_model.train()

    def load_state_dict(self, state_dict, strict=True):
        return self.net.load_state_dict(state_dict, strict)

    def state_dict(self):
        return self.net.state_dict()
