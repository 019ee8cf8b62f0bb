"""Box-vote decoding on the GPU: drop-ins for the reference's CPU torch loops.

NMS_clustering / mask_NMS keep the reference's signatures and return values
(/root/reference/models/iou_nms.py:68-105,130-144); detection2mask follows
/root/reference/models/detection_net.py:369-488 but keeps heat-maps and masks on the device,
bit-packed, until the final projection.
"""
import numpy as np
import torch

from . import ops

DECODE_THREADS = 4      # host threads (one CUDA stream each) that decode the scenes of a batch side by side
_DECODE_STREAMS = {}


def _decode_streams(dev, n):
    pool = _DECODE_STREAMS.setdefault(dev, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def NMS_clustering(boxes, cluster_th=0.5, get_heatmaps=True):
    """boxes f32[M,7] = (score, min, max). Returns (representatives i64[K], clusters: list of i64 tensors in
    descending-score order, heatmaps f32[K,M]) like the reference; tensors stay on boxes' CUDA device."""
    assert boxes.shape[1] == 7 and boxes.dim() == 2
    assert 0 < cluster_th < 1
    boxes = boxes.contiguous().float()
    reps, cluster_of, heat = ops.aabb_nms(boxes, cluster_th, want_heatmaps=get_heatmaps)
    # members of each cluster in descending-score order (ties: lower index): ONE sort of the M boxes by
    # (cluster, -score, index) and one split (a single host sync for the K sizes), instead of K boolean selections
    m = boxes.shape[0]
    cl = cluster_of.long()
    by_score = torch.argsort(-boxes[:, 0], stable=True)
    by_cluster = by_score[torch.sort(cl[by_score], stable=True)[1]]
    sizes = torch.bincount(cl, minlength=len(reps)).tolist() if m else []
    clusters = list(torch.split(by_cluster, sizes))
    if get_heatmaps:
        return reps, clusters, heat
    return reps, clusters


def mask_NMS(sorted_masks, cluster_th=0.5, allow_empty=False):
    """sorted_masks bool[K,N] (or bit-packed int32[K,words]); returns (kept indices i64[K'], None)."""
    packed = sorted_masks if sorted_masks.dtype == torch.int32 else ops.pack_masks_torch(sorted_masks)
    keep = ops.mask_nms(packed, cluster_th)
    return torch.nonzero(keep).flatten(), None


def to_bbs_min_max(locations, offsets, bounds, scores=None):
    """/root/reference/utils/util.py:46-64."""
    centers = offsets + locations
    bbs = torch.cat([centers - bounds, centers + bounds], dim=1)
    return torch.cat((scores, bbs), dim=1) if scores is not None else bbs


def detection2mask(net, batch, pred, cfg, mode="eval", score_filtering=True, cluster_th=0.3, score_th=0.3,
                   mask_bin_th=0.3, mask_nms_th=0.3):
    """pred: dict head -> f32 tensor (any device). Returns {scene name: {conf, label_id, mask, ...}}.

    Two semantics sources, like the reference (models/detection_net.py:378-415):
      * per-superpoint head (`cfg.mlp_semantics`; ScanNet, ARKitScenes): class ids through semantic_valid_class_ids,
        foreground per superpoint, mask-NMS on;
      * per-voxel head (`cfg.mlp_per_vox_semantics`, net.requires_voxel_outputs; S3DIS): class INDICES, every superpoint
        takes the mode of its voxels' labels (torch.mode: lowest label on ties), foreground from those, NO mask-NMS
        (:449-451). As in the reference this branch takes the per-voxel predictions of the whole batch for every
        scene, i.e. it is meant for batch size 1 (models/evaluation.py:70-91 decodes scene by scene)."""
    dev = torch.device(net.device)
    P = {k: v.to(dev).float() for k, v in pred.items()}
    loc = batch["input_location"].to(dev)
    pred_bbs = to_bbs_min_max(loc, P[cfg.mlp_offsets], P[cfg.mlp_bounds], torch.sigmoid(P[cfg.mlp_bb_scores]))
    per_vox = cfg.mlp_per_vox_semantics in cfg.network_heads
    voxel_outputs = bool(getattr(net, "requires_voxel_outputs", per_vox))
    if per_vox:
        pred_sem = torch.argmax(P[cfg.mlp_per_vox_semantics], 1)
        n_lab = P[cfg.mlp_per_vox_semantics].shape[1]
    else:
        class_ids = torch.as_tensor(net.semantic_valid_class_ids, device=dev).long()
        pred_sem = class_ids[torch.argmax(P[cfg.mlp_semantics], 1)]
        n_lab = int(class_ids.max().item()) + 1
    # rows of each scene: the collate function concatenates the scenes in order (models/dataloader.py:946-995), so a
    # scene is a contiguous range of superpoints and a slice replaces a boolean selection (each of which reads its size
    # back from the device); any other batch order falls back to the selections
    bid_src = batch["batch_ids"]
    n_scenes = len(batch["scene"])
    ranges = None
    if bid_src.numel() > 0 and bool((bid_src[1:] >= bid_src[:-1]).all()):
        counts = torch.bincount(bid_src.long().reshape(-1), minlength=n_scenes).tolist()
        ranges, a = [], 0
        for c in counts[:n_scenes]:
            ranges.append(slice(a, a + c))
            a += c
    batch_ids = batch["batch_ids"].to(dev)

    def decode_scene(scene_idx):
        scene_mask = ranges[scene_idx] if ranges is not None else batch_ids == scene_idx
        seg2vox = batch["seg2vox"][scene_idx].to(dev).long().contiguous()
        n_vox = seg2vox.shape[0]
        if voxel_outputs:
            # per-segment majority vote over the scene's voxels (segments are numbered by np.unique, :399-410)
            segments = torch.as_tensor(batch["vox_segments"][scene_idx], device=dev).long()
            uniq, seg_rank = torch.unique(segments, return_inverse=True)
            sem_vox = pred_sem.to(torch.int32).contiguous()
            seg_sem, _ = ops.segment_label_vote(seg_rank.contiguous(), sem_vox, int(uniq.shape[0]), n_lab)
            scene_fg = net.is_foreground(seg_sem.long())
        else:
            scene_sem = pred_sem[scene_mask]
            scene_fg = net.is_foreground(scene_sem)
            sem_vox = scene_sem[seg2vox].to(torch.int32).contiguous()
        scene_fg = torch.as_tensor(scene_fg, device=dev).bool()
        scene_bbs = pred_bbs[scene_mask][scene_fg].contiguous()
        if scene_bbs.shape[0] == 0:
            return {"conf": torch.zeros(0), "label_id": np.zeros(0, dtype="int32"),
                    "mask": torch.zeros((0, n_vox), dtype=torch.bool)}
        reps, cluster_of, heat = ops.aabb_nms(scene_bbs, cluster_th)
        scores = scene_bbs[reps][:, 0]
        if score_filtering:
            sel = scores > score_th
            heat, scores, reps = heat[sel].contiguous(), scores[sel], reps[sel]
        # rank of every foreground superpoint among the foreground ones (-1 for the others), without a size read-back
        fg_rank = torch.where(scene_fg, torch.cumsum(scene_fg, 0) - 1, -1).to(torch.int32)
        packed = ops.heatmap_project(heat, fg_rank, seg2vox, mask_bin_th)
        if not voxel_outputs:
            keep = ops.mask_nms(packed, mask_nms_th)
            packed, scores, reps = packed[keep].contiguous(), scores[keep], reps[keep]
        # per-instance majority label (np.bincount + argmax, detection_net.py:461-466) on the bit-packed masks
        labels, _ = ops.mask_label_vote(packed, sem_vox, n_vox, n_lab)
        labels = labels.cpu().numpy().astype("int32")
        masks = ops.unpack_masks(packed, n_vox)
        res = {"conf": scores.cpu(), "label_id": labels}
        if mode == "eval" and "vox2point" in batch:
            res["mask"] = masks[:, batch["vox2point"][scene_idx].to(dev)].cpu()
        else:
            res["mask"] = masks.cpu()
            res["cluster_representatives"] = reps.cpu()
            res["bbs"] = scene_bbs[reps].cpu()
            res["pred_fg"] = scene_fg.cpu()
        return res

    # The scenes of a batch are independent and each one's chain is a dozen small kernels around half a dozen reads of
    # a size from the device (boolean selections, the number of clusters): decoded one after the other the GPU mostly
    # waits for the host's round trips. With several scenes each one is decoded by its own host thread on its own
    # stream, so the round trips of one scene overlap the kernels of the others. Results are identical.
    names = [scene["name"] for scene in batch["scene"]]
    workers = min(len(names), DECODE_THREADS) if dev.type == "cuda" else 1
    if workers <= 1:
        return {name: decode_scene(i) for i, name in enumerate(names)}
    main = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main)              # the predictions above were produced on the caller's stream
    streams = _decode_streams(dev, workers)
    grad_mode = torch.is_grad_enabled()      # thread-local in torch: the workers take over the caller's mode

    def work(w):
        out = []
        with torch.set_grad_enabled(grad_mode), torch.cuda.device(dev), torch.cuda.stream(streams[w]):
            streams[w].wait_event(ready)
            for i in range(w, len(names), workers):
                out.append((i, decode_scene(i)))
            done = torch.cuda.Event()
            done.record(streams[w])
        return out, done

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as pool:
        parts = list(pool.map(work, range(workers)))
    results = [None] * len(names)
    for out, done in parts:
        main.wait_event(done)       # the scene streams read tensors of the caller's stream: order their release after that
        for i, res in out:
            results[i] = res
    return dict(zip(names, results))
