"""TEST INFRASTRUCTURE (CPU oracle) - weak-supervision label association, SURVEY.md section 8(f) row 2.

Restates /root/reference/models/dataloader.py:203-314 (`ScanNet.approx_association`): every point is tested against the
(foreground, optionally dropped-out / noised) instance boxes; a point in exactly one box takes that instance, a point in
none is background (-1), a point in several is undecided (-2) or takes the smallest box (`smallest_bb_heuristic`);
the per-point result is returned as is (`point_association`), majority-pooled per superpoint (`majority_vote`) or decided
per superpoint from its least-covered point (default). Vectorised numpy instead of the reference's Python loops; the
arithmetic (float64 comparisons, float64 box volumes, first-index tie breaks) is the reference's.

Pinned: tests/golden/label_assoc.npz holds the outputs of the reference's own function run on seeded synthetic scenes
(oracle/make_golden.py:golden_label_assoc); tests/test_oracle.py compares this restatement with them.
Only tests/, __graft_entry__.smoke() and bench.py's CPU leg may import this module."""
import numpy as np


def prepare_boxes(labels, scene_name, dropout_boxes=0.0, noisy_boxes=0.0):
    """dataloader.py:207-234: foreground boxes (semantics > 2 and != 22), optional per-scene-seeded dropout and corner
    noise, 5 mm padding. -> (min_corner f64[B,3], max_corner f64[B,3], instance_ids[B], volume f64[B])."""
    semantics = np.asarray(labels["per_instance_semantics"])
    scene_fg = (semantics > 2) & (semantics != 22)
    if dropout_boxes:
        rng = np.random.default_rng(seed=abs(int(scene_name, 36)))
        dropout_mask = rng.binomial(1, 1 - dropout_boxes, scene_fg.sum()) != 0
        scene_fg[scene_fg] = dropout_mask
    centers = np.asarray(labels["per_instance_bb_centers"])[scene_fg]
    bounds = np.asarray(labels["per_instance_bb_bounds"])[scene_fg] + 0.005
    min_corner = centers - bounds
    max_corner = centers + bounds
    instance_ids = np.asarray(labels["unique_instances"])[scene_fg]
    if noisy_boxes:
        rng = np.random.default_rng(seed=abs(int(scene_name, 36)))
        min_corner = min_corner + rng.normal(loc=0, scale=noisy_boxes / 2, size=min_corner.shape)
        max_corner = max_corner + rng.normal(loc=0, scale=noisy_boxes / 2, size=max_corner.shape)
    volume = np.prod(2 * bounds, axis=1)
    return min_corner, max_corner, instance_ids, volume


def point_boxes(positions, min_corner, max_corner, volume):
    """-> (num_boxes int64[N], first_box int64[N] (lowest box index containing the point, -1 if none), smallest_box
    int64[N] (containing box of least volume, ties: lowest index, -1 if none))."""
    pos = np.asarray(positions)
    n, b = len(pos), len(min_corner)
    if b == 0:
        return np.zeros(n, np.int64), np.full(n, -1, np.int64), np.full(n, -1, np.int64)
    occ = np.all(pos[None] >= min_corner[:, None], axis=-1) & np.all(pos[None] <= max_corner[:, None], axis=-1)   # [B, N]
    num = occ.sum(0).astype(np.int64)
    first = np.where(num > 0, occ.argmax(0), -1).astype(np.int64)
    vol = np.where(occ, volume[:, None], np.inf)
    smallest = np.where(num > 0, vol.argmin(0), -1).astype(np.int64)
    return num, first, smallest


def approx_association(positions, segments, unique_segs, min_corner, max_corner, instance_ids, volume,
                       point_association=False, majority_vote=False, smallest_bb_heuristic=False):
    """dataloader.py:236-314 on prepared boxes. -> (inst_per_point int64[N], inst_per_seg int64[S] or None)."""
    num, first, smallest = point_boxes(positions, min_corner, max_corner, volume)
    ids = np.asarray(instance_ids).astype(np.int64)
    segments = np.asarray(segments)
    n = len(num)
    if point_association or majority_vote:
        inst = np.full(n, -1, np.int64)
        one = num == 1
        inst[one] = ids[first[one]]
        many = num > 1
        inst[many] = ids[smallest[many]] if smallest_bb_heuristic else -2
        if point_association:
            return inst, None
        pooled = np.full(n, -2, np.int64)
        per_seg = np.full(len(unique_segs), -2, np.int64)
        for i, seg_id in enumerate(unique_segs):
            m = segments == seg_id
            vals, counts = np.unique(inst[m], return_counts=True)       # scipy.stats.mode: smallest of the most common
            v = vals[counts.argmax()]
            pooled[m] = v
            per_seg[i] = v
        return pooled, per_seg
    pooled = np.full(n, -2, np.int64)
    per_seg = np.full(len(unique_segs), -2, np.int64)
    for i, seg_id in enumerate(unique_segs):
        idx = np.nonzero(segments == seg_id)[0]
        nb = num[idx]
        mn = nb.min()
        if mn == 1:
            p = idx[np.nonzero(nb == 1)[0][0]]
            per_seg[i] = ids[first[p]]
        elif mn == 0:
            per_seg[i] = -1
        elif smallest_bb_heuristic:
            p = idx[nb.argmin()]
            per_seg[i] = ids[smallest[p]]
        pooled[idx] = per_seg[i]
    return pooled, per_seg


def synthetic_case(seed, n_points=8000, n_inst=14):
    """A seeded scene for the goldens and the tests: points on three planes, superpoints = 0.3 m cells of each plane
    (ids with gaps, not sorted by position), boxes that overlap each other and cut through superpoints.
    -> (labels dict, scene dict, unique_segs)."""
    rng = np.random.default_rng(seed)
    plane = rng.integers(0, 3, n_points)
    pos = rng.uniform(0, 4, (n_points, 3))
    pos[np.arange(n_points), plane] = np.array([0.0, 0.05, 3.9])[plane] + rng.normal(0, 0.004, n_points)
    uv = np.stack([np.delete(p, ax) for p, ax in zip(pos, plane)])
    cell = (plane * 14 + np.floor(uv[:, 0] / 0.3).astype(int)) * 14 + np.floor(uv[:, 1] / 0.3).astype(int)
    relabel = rng.permutation(4 * (cell.max() + 1))
    segments = relabel[cell].astype(np.int64)
    unique_segs = np.unique(segments)
    centers = rng.uniform(0.3, 3.7, (n_inst, 3))
    centers[: n_inst // 2, rng.integers(0, 3)] = 0.1          # half of the boxes sit on a plane
    bounds = rng.uniform(0.15, 0.8, (n_inst, 3))
    semantics = rng.choice([1, 2, 3, 5, 8, 22, 0, 12, 33, 7, 9], n_inst)
    labels = {"per_instance_semantics": semantics, "per_instance_bb_centers": centers, "per_instance_bb_bounds": bounds,
              "unique_instances": np.arange(1, n_inst + 1) * 3}
    scene = {"name": "scene%04d" % seed, "positions": pos, "segments": segments}
    return labels, scene, unique_segs
