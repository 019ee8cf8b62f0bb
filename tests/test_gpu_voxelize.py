"""GPU voxelisation + collate (box2mask_b200/voxelize.py -> b2m_voxel_coords, b2m_downsample_coords,
b2m_kernel_map_submanifold, b2m_nearest_point) against the CPU restatement of the reference's data loader
(oracle/voxelize.py: numpy round / unique, scikit-learn ball tree). Index maps must be identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from box2mask_b200 import voxelize as vz  # noqa: E402
from oracle import voxelize as ovz  # noqa: E402

DEV = "cuda"


def _cloud(seed, p, lo, hi, n_seg):
    rng = np.random.default_rng(seed)
    # points on a few planes plus uniform clutter: many voxels with several points, some with one far from the centre
    a = rng.uniform(lo, hi, (p // 2, 3))
    b = rng.uniform(lo, hi, (p - p // 2, 3))
    b[:, 2] = np.round(b[:, 2] * 2) / 2 + rng.normal(0, 0.002, len(b))
    pos = np.concatenate([a, b], 0)
    col = rng.normal(size=(p, 3)).astype(np.float32)
    nor = rng.normal(size=(p, 3)).astype(np.float32)
    seg = (rng.integers(0, n_seg, p) * 7 + 3).astype(np.int64)       # sparse, unsorted segment ids
    return pos, col, nor, seg


@pytest.mark.parametrize("seed,p,lo,hi,voxel", [(1, 40000, -0.4, 1.6, 0.02), (2, 60000, 0.1, 2.0, 0.04), (3, 500, -1.0, -0.2, 0.05)])
def test_voxelize_scene_matches_reference_restatement(seed, p, lo, hi, voxel):
    pos, col, nor, seg = _cloud(seed, p, lo, hi, 60)
    ref = ovz.voxelize_scene(pos, col, nor, seg, voxel)
    got = vz.voxelize_scene(torch.from_numpy(pos).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(nor).to(DEV),
                            torch.from_numpy(seg).to(DEV), voxel)
    assert int(got["status"].item()) == 0
    assert np.array_equal(got["vox_coords"].cpu().numpy(), ref["vox_coords"].astype(np.int32))
    assert np.array_equal(got["vox2point"].cpu().numpy(), ref["vox2point"])
    assert np.array_equal(got["point2vox"].cpu().numpy(), ref["point2vox"])
    assert np.array_equal(got["vox_features"].cpu().numpy(), ref["vox_features"])
    assert np.array_equal(got["vox_segments"].cpu().numpy(), ref["vox_segments"])
    assert np.array_equal(got["seg2vox"].cpu().numpy(), ref["seg2vox"])
    assert np.array_equal(got["seg2point"].cpu().numpy(), ref["seg2point"])
    assert np.array_equal(got["unique_vox_segments"].cpu().numpy(), ref["unique_vox_segments"])
    assert np.array_equal(got["vox_world_coords"].cpu().numpy(), ref["vox_world_coords"])
    # segment means: fp64 sums in a different order, then the collate casts to fp32
    assert np.allclose(got["input_location"].cpu().numpy(), ref["input_location"], rtol=0, atol=1e-12)


def test_collate_matches_reference_restatement():
    items_ref, items = [], []
    for seed in (5, 6, 7):
        pos, col, nor, seg = _cloud(seed, 20000, 0.0, 1.5, 40)
        items_ref.append(ovz.voxelize_scene(pos, col, nor, seg, 0.02))
        items.append(vz.voxelize_scene(torch.from_numpy(pos).to(DEV), torch.from_numpy(col).to(DEV),
                                       torch.from_numpy(nor).to(DEV), torch.from_numpy(seg).to(DEV), 0.02))
    b = vz.collate_scenes(items)
    ref_coords = np.concatenate([np.concatenate([np.full((len(r["vox_coords"]), 1), i), r["vox_coords"]], 1)
                                 for i, r in enumerate(items_ref)], 0).astype(np.int32)
    assert np.array_equal(b["vox_coords"].cpu().numpy(), ref_coords)
    assert np.array_equal(b["pooling_ids"].cpu().numpy(), ovz.to_unique([r["vox_segments"] for r in items_ref]))
    assert b["num_segments"] == sum(len(r["unique_vox_segments"]) for r in items_ref)
    assert np.array_equal(b["vox_features"].cpu().numpy(), np.concatenate([r["vox_features"] for r in items_ref], 0))
    ref_loc = np.concatenate([r["input_location"] for r in items_ref], 0).astype(np.float32)
    assert np.allclose(b["input_location"].cpu().numpy(), ref_loc, rtol=0, atol=1e-6)
    assert np.array_equal(b["batch_ids"].cpu().numpy(),
                          np.concatenate([np.full(len(r["input_location"]), i) for i, r in enumerate(items_ref)]))
