"""GPU label association (csrc/assoc.cu, box2mask_b200/label_assoc.py) against the outputs of the reference's own
ScanNet.approx_association (tests/golden/label_assoc.npz) and against the oracle on larger random cases: exact integers."""
import os
import types

import numpy as np
import pytest
import torch

from box2mask_b200 import label_assoc as la
from oracle import label_assoc as ola

pytestmark = pytest.mark.gpu


def test_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "label_assoc.npz"))
    tags = sorted(k[:-4] for k in g.files if k.endswith("_cfg"))
    assert len(tags) == 7
    for tag in tags:
        seed, pa, mv, small, drop, noise = g[tag + "_cfg"]
        labels, scene, unique_segs = ola.synthetic_case(int(seed))
        cfg = types.SimpleNamespace(dropout_boxes=drop, noisy_boxes=noise, smallest_bb_heuristic=bool(small))
        per_point, per_seg = la.approx_association(labels, scene, cfg, bool(pa), bool(mv), unique_segs)
        assert np.array_equal(per_point, g[tag + "_per_point"]), tag
        if per_seg is not None:
            assert np.array_equal(per_seg, g[tag + "_per_seg"]), tag


@pytest.mark.parametrize("n_points,n_inst", [(250000, 60), (1000, 1), (5000, 700), (64, 0)])
def test_matches_oracle_large_and_edge_cases(n_points, n_inst):
    """ScanNet-size scene (250 k points), a single box, more boxes than one shared-memory pass holds, no boxes at all."""
    labels, scene, unique_segs = ola.synthetic_case(7, n_points=n_points, n_inst=max(n_inst, 1))
    if n_inst == 0:
        labels["per_instance_semantics"][:] = 1            # every box is a wall: nothing survives the foreground filter
    else:
        labels["per_instance_semantics"][:] = 5
    # drop a few superpoints from the list, as voxelisation does (their points must come back as -2)
    unique_segs = unique_segs[::3] if len(unique_segs) > 8 else unique_segs
    for pa, mv, small in [(False, False, False), (False, False, True), (True, False, True), (False, True, False), (False, True, True)]:
        cfg = types.SimpleNamespace(dropout_boxes=0.0, noisy_boxes=0.0, smallest_bb_heuristic=small)
        mn, mx, ids, vol = ola.prepare_boxes(labels, scene["name"])
        keep = np.isin(scene["segments"], unique_segs)
        ref_point, ref_seg = ola.approx_association(scene["positions"][keep], scene["segments"][keep], unique_segs, mn, mx, ids, vol,
                                                    pa, mv, small)
        per_point, per_seg = la.approx_association(labels, scene, cfg, pa, mv, unique_segs)
        if pa:
            full, _ = ola.approx_association(scene["positions"], scene["segments"], unique_segs, mn, mx, ids, vol, True, False, small)
            assert np.array_equal(per_point, full)
            continue
        assert np.array_equal(per_seg, ref_seg), (pa, mv, small)
        assert np.array_equal(per_point[keep], ref_point)
        assert (per_point[~keep] == -2).all()


def test_occupancy_counts_against_numpy():
    rng = np.random.default_rng(0)
    pos = rng.uniform(0, 1, (10000, 3))
    mn = rng.uniform(0, 0.8, (40, 3))
    mx = mn + rng.uniform(0.05, 0.5, (40, 3))
    vol = np.prod(mx - mn, axis=1)
    num, first, smallest = la.point_box_occupancy(torch.as_tensor(pos, device="cuda"), torch.as_tensor(mn, device="cuda"),
                                                  torch.as_tensor(mx, device="cuda"), torch.as_tensor(vol, device="cuda"))
    rn, rf, rs = ola.point_boxes(pos, mn, mx, vol)
    assert np.array_equal(num.cpu().numpy(), rn) and np.array_equal(first.cpu().numpy(), rf) and np.array_equal(smallest.cpu().numpy(), rs)
