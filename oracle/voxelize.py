"""CPU restatement of the reference's per-scene voxelisation and collate.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/models/dataloader.py:61-124 (shift, scale, round, unique rows with inverse; one nearest scene
point per voxel centre through scikit-learn's ball tree; features / segments of that point; mean world position of the
voxels of every segment) and models/dataloader.py:946-995 with utils/util.py:123-130 (batched coordinates, dense
cross-scene segment ids). The reference needs the ScanNet files to run this code path, so there is no fixture:
"parity unpinned" for this next-row component; the restatement uses the same numpy / scikit-learn calls.
"""
import numpy as np
from sklearn.neighbors import NearestNeighbors


def voxelize_scene(positions, colors, normals, segments, voxel_size, use_normals=True):
    shift = min(0, np.min(positions))
    scaled = (positions - shift) / voxel_size                                   # dataloader.py:63-65
    vox_coords, vox2point = np.unique(np.round(scaled), axis=0, return_inverse=True)   # :67-68
    vox2point = vox2point.reshape(-1)
    tree = NearestNeighbors(n_neighbors=1, algorithm="ball_tree").fit(scaled)   # :75-77
    point2vox = tree.kneighbors(vox_coords, return_distance=False).reshape(-1)
    feats = np.concatenate([colors, normals], 1) if use_normals else colors     # :82-91
    world = vox_coords * voxel_size + shift                                     # :95
    vox_segments = segments[point2vox]
    uniq, seg2vox = np.unique(vox_segments, return_inverse=True)                # :110
    middle = np.stack([world[vox_segments == u].mean(axis=0) for u in uniq])    # :113-117
    return {"vox_coords": vox_coords, "vox2point": vox2point, "point2vox": point2vox, "vox_features": feats[point2vox],
            "vox_segments": vox_segments, "vox_world_coords": world, "seg2vox": seg2vox.reshape(-1),
            "unique_vox_segments": uniq, "seg2point": seg2vox.reshape(-1)[vox2point], "input_location": middle}


def to_unique(segments):
    """utils/util.py:123-130."""
    segs = [np.array(s, copy=True) for s in segments]
    for i in range(1, len(segs)):
        segs[i] += np.max(segs[i - 1]) + 1
    _, ids = np.unique(np.concatenate(segs, 0), return_inverse=True)
    return ids.reshape(-1)
