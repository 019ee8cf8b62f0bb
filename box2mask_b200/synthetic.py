"""Deterministic synthetic scenes of ScanNet / S3DIS / ARKitScenes shape (no real data offline).

A room (floor + 4 walls) with random axis-aligned furniture boxes, surfaces sampled densely with
Gaussian noise and voxelised exactly like the reference data loader
(/root/reference/models/dataloader.py:63-68: shift non-negative, divide by voxel size, round, unique),
so coordinates are unique, non-negative and lexicographically sorted. Batches follow the reference
collate (/root/reference/models/dataloader.py:946-995; ids by utils/util.py:123-130).
"""
import numpy as np
import torch


def make_scene(seed, scale=0.84, voxel_size=0.02, density=2.0e4, n_classes=20):
    """One scene. Returns a dict of numpy arrays:
    vox_coords int32[N,3], vox_features f32[N,6], vox_segments int64[N] (dense 0..S-1),
    input_location f32[S,3], gt_bb_offsets f32[S,3], gt_bb_bounds f32[S,3], gt_semantics int64[S],
    fg_instances bool[S]."""
    rng = np.random.default_rng(seed)
    L, W, H = 4.2 * scale, 3.4 * scale, 1.8
    pts, nrm = [], []

    def plane(origin, u, v, normal):
        area = np.linalg.norm(u) * np.linalg.norm(v)
        k = max(int(area * density), 1)
        a, b = rng.random(k), rng.random(k)
        p = origin[None] + a[:, None] * u[None] + b[:, None] * v[None]
        pts.append(p)
        nrm.append(np.repeat(np.asarray(normal, dtype=np.float64)[None], k, 0))

    o = np.zeros(3)
    ex, ey, ez = np.array([L, 0, 0.]), np.array([0, W, 0.]), np.array([0, 0, H])
    plane(o, ex, ey, (0, 0, 1))            # floor
    plane(o, ex, ez, (0, 1, 0))            # wall y=0
    plane(o + ey, ex, ez, (0, -1, 0))      # wall y=W
    plane(o, ey, ez, (1, 0, 0))            # wall x=0
    plane(o + ex, ey, ez, (-1, 0, 0))      # wall x=L
    boxes = []
    for _ in range(int(10 * scale * scale)):
        sx, sy, sz = rng.uniform(0.4, 1.6), rng.uniform(0.4, 0.9), rng.uniform(0.4, 1.0)
        cx, cy = rng.uniform(0, max(L - sx, 0.1)), rng.uniform(0, max(W - sy, 0.1))
        b0 = np.array([cx, cy, 0.0])
        bx, by, bz = np.array([sx, 0, 0.]), np.array([0, sy, 0.]), np.array([0, 0, sz])
        plane(b0 + bz, bx, by, (0, 0, 1))
        plane(b0, bx, bz, (0, -1, 0))
        plane(b0 + by, bx, bz, (0, 1, 0))
        plane(b0, by, bz, (-1, 0, 0))
        plane(b0 + bx, by, bz, (1, 0, 0))
        boxes.append((b0, b0 + np.array([sx, sy, sz])))
    p = np.concatenate(pts, 0)
    nr = np.concatenate(nrm, 0)
    p = p + rng.normal(0.0, 0.0015, p.shape)
    # reference voxelisation (models/dataloader.py:63-68)
    p = p - min(0.0, p.min())
    q = np.round(p / voxel_size)
    coords, first = np.unique(q, axis=0, return_index=True)
    coords = coords.astype(np.int32)
    n = len(coords)
    feats = np.concatenate([rng.normal(0, 1, (n, 3)), nr[first]], 1).astype(np.float32)
    # superpoints: coarse 10-voxel cell x dominant-normal bucket, made dense
    cell = coords // 10
    bucket = np.argmax(np.abs(nr[first]), 1)
    seg_key = ((cell[:, 0].astype(np.int64) * 4096 + cell[:, 1]) * 4096 + cell[:, 2]) * 4 + bucket
    _, segs = np.unique(seg_key, return_inverse=True)
    s = int(segs.max()) + 1
    # per-segment mean voxel location in metres (models/dataloader.py:110-120)
    cnt = np.bincount(segs, minlength=s).astype(np.float64)
    loc = np.stack([np.bincount(segs, weights=coords[:, d].astype(np.float64), minlength=s) / cnt for d in range(3)], 1)
    loc = (loc * voxel_size).astype(np.float32)
    fg = rng.random(s) < 0.4
    return {
        "vox_coords": coords,
        "vox_features": feats,
        "vox_segments": segs.astype(np.int64),
        "input_location": loc,
        "gt_bb_offsets": rng.uniform(-1, 1, (s, 3)).astype(np.float32),
        "gt_bb_bounds": rng.uniform(0.05, 1, (s, 3)).astype(np.float32),
        "gt_semantics": rng.integers(0, n_classes + 1, s).astype(np.int64),   # 0 = unlabeled (ignored)
        "fg_instances": fg,
        "gt_boxes": boxes,
    }


def label_maps(n_classes=20):
    """ScanNet-like label tables (dataprocessing/scannet.py:114-118): valid class ids 1..n, id 0 unlabeled -> -100,
    ids 1 and 2 (wall, floor) are background."""
    valid = torch.arange(1, n_classes + 1)
    id2idx = torch.cat([torch.tensor([-100]), torch.arange(n_classes)])

    def is_foreground(sem_ids):
        return sem_ids > 2
    return valid, id2idx, is_foreground


def batched_coordinates(coords_list, dtype=torch.int32):
    """ME.utils.batched_coordinates as used at /root/reference/models/dataloader.py:966:
    list of [n_i, 3] -> int32 [sum n_i, 4] with the batch index in column 0. CPU only."""
    out = []
    for b, c in enumerate(coords_list):
        c = torch.as_tensor(np.asarray(c)) if not torch.is_tensor(c) else c
        c = c.to(dtype)
        out.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=dtype), c], 1))
    return torch.cat(out, 0) if out else torch.zeros((0, 4), dtype=dtype)


def to_unique(segments):
    """Dense cross-scene superpoint ids (/root/reference/utils/util.py:123-130)."""
    segs = [np.array(s, copy=True) for s in segments]
    for i in range(1, len(segs)):
        segs[i] += np.max(segs[i - 1]) + 1
    allsegs = np.concatenate(segs, 0)
    _, ids = np.unique(allsegs, return_inverse=True)
    return torch.from_numpy(ids).long()


def collate(scenes):
    """Batch dict with the reference collate_fn's keys (models/dataloader.py:946-995)."""
    ret = {
        "vox_coords": batched_coordinates([s["vox_coords"] for s in scenes]),
        "vox_features": torch.from_numpy(np.concatenate([s["vox_features"] for s in scenes], 0)).float(),
        "batch_ids": torch.from_numpy(np.concatenate(
            [np.full(len(s["input_location"]), b) for b, s in enumerate(scenes)], 0)).long(),
        "input_location": torch.from_numpy(np.concatenate([s["input_location"] for s in scenes], 0)).float(),
        "pooling_ids": to_unique([s["vox_segments"] for s in scenes]),
        "gt_bb_bounds": torch.from_numpy(np.concatenate([s["gt_bb_bounds"] for s in scenes], 0)).float(),
        "gt_bb_offsets": torch.from_numpy(np.concatenate([s["gt_bb_offsets"] for s in scenes], 0)).float(),
        "gt_semantics": torch.from_numpy(np.concatenate([s["gt_semantics"] for s in scenes], 0)).long(),
        "fg_instances": torch.from_numpy(np.concatenate([s["fg_instances"] for s in scenes], 0)).bool(),
        "seg2vox": [torch.from_numpy(s["vox_segments"]).long() for s in scenes],
        "scene": [{"name": "synthetic_%d" % b} for b in range(len(scenes))],
    }
    # optional key (not in the reference's collate_fn): the rows of the foreground superpoints, so that the losses gather
    # them without reading the size of a boolean-mask selection back from the device (Model.compute_loss_detection)
    ret["fg_index"] = torch.nonzero(ret["fg_instances"]).reshape(-1)
    return ret


def make_batch(n_scenes, seed=0, scale=0.84, voxel_size=0.02, n_classes=20, density=2.0e4):
    return collate([make_scene(1000 * seed + i, scale, voxel_size, density, n_classes) for i in range(n_scenes)])


def make_boxes(m=2000, n_centres=150, seed=0):
    """Direct synthetic NMS input: m boxes jittered around n_centres, f32[m,7]."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0, 4, (n_centres, 3))
    sizes = rng.uniform(0.2, 0.8, (n_centres, 3))
    which = rng.integers(0, n_centres, m)
    c = centres[which] + rng.normal(0, 0.05, (m, 3))
    h = sizes[which] * rng.uniform(0.8, 1.2, (m, 3))
    score = rng.random((m, 1))
    return torch.from_numpy(np.concatenate([score, c - h, c + h], 1).astype(np.float32))
