"""Weak-supervision label association on the GPU: drop-in for
`ScanNet.approx_association` (/root/reference/models/dataloader.py:203-314).

The reference walks every point and every superpoint in Python (a `[boxes x points]` occupancy matrix, a list of
`np.argwhere` per point, a boolean mask per superpoint). Here the box preparation (a few dozen boxes: foreground filter,
per-scene-seeded dropout / corner noise, 5 mm padding) stays on the host exactly as in the reference, and the per-point /
per-superpoint work runs in libb2m.so (csrc/assoc.cu). Results are identical integers (tests/test_gpu_label_assoc.py
against the outputs of the reference's own function)."""
import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def prepare_boxes(labels, scene_name, dropout_boxes=0.0, noisy_boxes=0.0, ret=None):
    """dataloader.py:207-234. -> (min_corner f64[B,3], max_corner f64[B,3], instance_ids int64[B], volume f64[B])."""
    semantics = np.asarray(labels["per_instance_semantics"])
    scene_fg = (semantics > 2) & (semantics != 22)          # no boxes from walls / floor / ceiling / unlabeled
    if dropout_boxes:
        rng = np.random.default_rng(seed=abs(int(scene_name, 36)))      # the same instances per scene every time
        scene_fg[scene_fg] = rng.binomial(1, 1 - dropout_boxes, scene_fg.sum()) != 0
    centers = np.asarray(labels["per_instance_bb_centers"], dtype=np.float64)[scene_fg]
    bounds = np.asarray(labels["per_instance_bb_bounds"], dtype=np.float64)[scene_fg] + 0.005
    min_corner, max_corner = centers - bounds, centers + bounds
    instance_ids = np.asarray(labels["unique_instances"])[scene_fg].astype(np.int64)
    if noisy_boxes:
        rng = np.random.default_rng(seed=abs(int(scene_name, 36)))
        min_corner = min_corner + rng.normal(loc=0, scale=noisy_boxes / 2, size=min_corner.shape)
        max_corner = max_corner + rng.normal(loc=0, scale=noisy_boxes / 2, size=max_corner.shape)
        if ret is not None:
            ret["noisy_bbs"] = min_corner, max_corner
    return min_corner, max_corner, instance_ids, np.prod(2 * bounds, axis=1)


def point_box_occupancy(positions, min_corner, max_corner, volume):
    """positions f64[N,3] (cuda) -> (num, first, smallest) int32[N] (cuda)."""
    lib = _lib.load()
    n, b = positions.shape[0], min_corner.shape[0]
    num = torch.empty(n, dtype=torch.int32, device=positions.device)
    first, smallest = torch.empty_like(num), torch.empty_like(num)
    check(lib.b2m_point_box_occupancy(ptr(positions), n, ptr(min_corner), ptr(max_corner), ptr(volume), b, ptr(num), ptr(first),
                                      ptr(smallest), stream_ptr()), "point_box_occupancy")
    return num, first, smallest


def approx_association(labels, scene, cfg, point_association, majority_vote, unique_segs, ret=None, device="cuda"):
    """Same arguments and return values as the reference method (the dataset object's `cfg` is passed explicitly):
    -> (inst_per_point int64[N], inst_per_seg int64[S] or None) as numpy arrays; -1 background, -2 unknown."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.B2MError("label association runs on a CUDA device (no CPU fallback in the product path)")
    mn, mx, ids, vol = prepare_boxes(labels, scene["name"], getattr(cfg, "dropout_boxes", 0.0), getattr(cfg, "noisy_boxes", 0.0), ret)
    heur = 1 if getattr(cfg, "smallest_bb_heuristic", False) else 0
    pos = torch.as_tensor(np.ascontiguousarray(scene["positions"], dtype=np.float64), device=dev)
    n, b = pos.shape[0], len(ids)
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)     # noqa: E731
    mn_d, mx_d, vol_d = t(mn, torch.float64), t(mx, torch.float64), t(vol, torch.float64)
    num, first, smallest = point_box_occupancy(pos, mn_d, mx_d, vol_d)
    ids_d = t(ids, torch.int64)
    if point_association:
        inst = torch.empty(n, dtype=torch.int64, device=dev)
        check(lib.b2m_point_instances(ptr(num), ptr(first), ptr(smallest), ptr(ids_d), n, heur, ptr(inst), stream_ptr()),
              "point_instances")
        return inst.cpu().numpy(), None
    useg = torch.as_tensor(np.ascontiguousarray(unique_segs), dtype=torch.int64, device=dev)       # sorted (np.unique)
    segs = torch.as_tensor(np.ascontiguousarray(scene["segments"]), dtype=torch.int64, device=dev)
    s = useg.shape[0]
    rank = torch.searchsorted(useg, segs).clamp(max=max(s - 1, 0))
    seg_rank = torch.where(useg[rank] == segs, rank, torch.full_like(rank, -1)).to(torch.int32) if s else \
        torch.full((n,), -1, dtype=torch.int32, device=dev)
    order = np.argsort(ids, kind="stable")
    id_rank = np.empty(b, dtype=np.int32)
    id_rank[order] = np.arange(b, dtype=np.int32)
    ws_bytes = int(lib.b2m_segment_association_workspace_bytes(s, b, 1 if majority_vote else 0))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    per_seg = torch.empty(s, dtype=torch.int64, device=dev)
    per_point = torch.empty(n, dtype=torch.int64, device=dev)
    sorted_ids_d, id_rank_d = t(ids[order], torch.int64), t(id_rank, torch.int32)     # (kept alive across the launch)
    check(lib.b2m_segment_association(ptr(num), ptr(first), ptr(smallest), ptr(seg_rank), n, s, ptr(ids_d),
                                      ptr(sorted_ids_d), ptr(id_rank_d), b, 1 if majority_vote else 0,
                                      heur, ptr(per_seg), ptr(per_point), ptr(ws), ws_bytes, stream_ptr()), "segment_association")
    return per_point.cpu().numpy(), per_seg.cpu().numpy()
