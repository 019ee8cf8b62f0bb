"""Scene voxelisation and batch collation on the GPU: the producer of the hot path's inputs.

Restates, for device-resident points, what the reference does per scene on the CPU with numpy / scikit-learn
(/root/reference/models/dataloader.py:61-124) and per batch in `collate_fn` (models/dataloader.py:946-995;
ids by /root/reference/utils/util.py:123-130). Every output equals the reference's: the voxel set is the sorted
unique set of rounded coordinates, `vox2point` / `point2vox` / `seg2vox` are the same index maps, and the nearest
point of a voxel centre is found exactly (csrc/voxel.cu). There is no CPU fallback.
"""
import torch

from . import ops


def voxelize_scene(positions, colors, normals, segments, voxel_size, use_normals=True):
    """positions f64[P,3] (scene points), colors f32[P,3], normals f32[P,3], segments int64[P] (superpoint of every
    point), all on one CUDA device. Returns a dict with the reference's per-scene keys:
    vox_coords int32[N,3] (sorted lexicographically, unique), vox2point int64[P], point2vox int64[N],
    vox_features f32[N,6|3], vox_segments int64[N], vox_world_coords f64[N,3], seg2vox int64[N] (dense segment rank of
    every voxel), unique_vox_segments int64[S], seg2point int64[P], input_location f64[S,3] (mean voxel position of a
    segment in world space)."""
    if not positions.is_cuda:
        raise ops._lib.B2MError("voxelize_scene runs on a CUDA device only (no CPU fallback)")
    pos = positions.to(torch.float64).contiguous()
    min_pos = pos.min().reshape(1)                                       # stays on the device
    pcoords, status = ops.voxel_coords(pos, min_pos, voxel_size)
    vox4, vox2point = ops.downsample_coords(pcoords, 1)                  # sorted unique + inverse (one host sync: N)
    n = vox4.shape[0]
    vox2point = vox2point.long()
    # points grouped by voxel (CSR) and the 27-neighbourhood of every voxel
    point_order = torch.argsort(vox2point, stable=True)
    counts = torch.bincount(vox2point, minlength=n)
    start = torch.zeros(n + 1, dtype=torch.int64, device=pos.device)
    torch.cumsum(counts, 0, out=start[1:])
    nbr = ops.kernel_map_submanifold(vox4, 1, 3, ops.hash_build(vox4))
    point2vox = ops.nearest_point(pos, min_pos, voxel_size, vox4, nbr, start, point_order)
    feats = torch.cat([colors, normals], 1) if use_normals else colors
    vox_coords = vox4[:, 1:].contiguous()
    shift = torch.clamp(min_pos, max=0.0)
    world = vox_coords.double() * voxel_size + shift
    vox_segments = segments[point2vox]
    uniq, seg2vox = torch.unique(vox_segments, return_inverse=True)
    s = uniq.shape[0]
    sums = torch.zeros((s, 3), dtype=torch.float64, device=pos.device).index_add_(0, seg2vox, world)
    cnt = torch.bincount(seg2vox, minlength=s).double()
    return {
        "vox_coords": vox_coords, "vox2point": vox2point, "point2vox": point2vox,
        "vox_features": feats[point2vox].float(), "vox_segments": vox_segments, "vox_world_coords": world,
        "seg2vox": seg2vox, "unique_vox_segments": uniq, "seg2point": seg2vox[vox2point],
        "input_location": sums / cnt[:, None], "status": status,
    }


def collate_scenes(items):
    """The tensor part of the reference collate_fn for a list of voxelize_scene results: batched coordinates with the
    batch index in column 0 (ME.utils.batched_coordinates), concatenated features / locations, `batch_ids`, and the
    dense cross-scene superpoint ids (`to_unique`: scene-major, ascending segment id within a scene)."""
    dev = items[0]["vox_coords"].device
    coords, ids, off = [], [], 0
    for b, it in enumerate(items):
        c = it["vox_coords"]
        coords.append(torch.cat([torch.full((c.shape[0], 1), b, dtype=torch.int32, device=dev), c.int()], 1))
        ids.append(it["seg2vox"] + off)              # seg2vox is already the dense rank of the ascending segment ids
        off += int(it["unique_vox_segments"].shape[0])
    return {
        "vox_coords": torch.cat(coords, 0),
        "vox_features": torch.cat([it["vox_features"] for it in items], 0).float(),
        "pooling_ids": torch.cat(ids, 0),
        "input_location": torch.cat([it["input_location"] for it in items], 0).float(),
        "batch_ids": torch.cat([torch.full((it["input_location"].shape[0],), b, dtype=torch.int64, device=dev)
                                for b, it in enumerate(items)], 0),
        "seg2vox": [it["seg2vox"] for it in items],
        "vox2point": [it["vox2point"] for it in items],
        "num_segments": off,
    }
