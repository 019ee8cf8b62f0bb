"""CPU tests of the oracle: against the reference-generated golden vectors (NMS), against hand-countable
kernel maps, and against the independent dense-convolution oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import nms as onms
from oracle import sparse_ops as so


# ------------------------------------------------------------------------------------------------
# whole-scene decode: the oracle against the fixture produced by the reference's own detection2mask
# (models/detection_net.py:369-488, run unmodified in the build container by oracle/make_golden.py)
# ------------------------------------------------------------------------------------------------
def test_decode_scene_matches_reference_detection2mask(golden_dir):
    from box2mask_b200.selection_net import default_config
    from box2mask_b200.synthetic import label_maps
    from oracle.make_golden import decode_inputs
    g = np.load(os.path.join(golden_dir, "decode_small.npz"))
    cfg = default_config()
    valid, _, is_fg = label_maps(20)
    batch, pred = decode_inputs()
    boxes = onms.to_boxes(batch["input_location"], pred["mlp_offsets"], pred["mlp_bounds"],
                          torch.sigmoid(pred["mlp_bb_scores"]))
    sem = valid[torch.argmax(pred["mlp_semantics"], 1)].long()
    for b, scene in enumerate(batch["scene"]):
        m = batch["batch_ids"] == b
        fg = is_fg(sem[m])
        r = onms.decode_scene(boxes[m][fg], fg, batch["seg2vox"][b], sem[m], *cfg.eval_ths)
        name = scene["name"]
        assert np.array_equal(r["conf"].numpy(), g["train_%s_conf" % name])
        assert np.array_equal(r["label_id"], g["train_%s_label_id" % name])
        assert np.array_equal(r["representatives"].numpy(), g["train_%s_reps" % name])
        shape = tuple(g["train_%s_mask_shape" % name])
        ref_mask = np.unpackbits(g["train_%s_mask" % name], axis=1)[:, :shape[1]].astype(bool)
        assert np.array_equal(r["mask"].numpy(), ref_mask)
        shape = tuple(g["eval_%s_mask_shape" % name])
        ref_pts = np.unpackbits(g["eval_%s_mask" % name], axis=1)[:, :shape[1]].astype(bool)
        assert np.array_equal(r["mask"][:, batch["vox2point"][b]].numpy(), ref_pts)


# ------------------------------------------------------------------------------------------------
# NMS: bit-exact against fixtures produced by the reference's own models/iou_nms.py
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["nms_small", "nms_medium", "nms_lowth"])
def test_nms_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    boxes = torch.from_numpy(g["boxes"])
    reps, clusters, heat = onms.nms_clustering(boxes, float(g["th"]))
    assert np.array_equal(reps.numpy(), g["reps"])
    assert np.array_equal(heat.numpy(), g["heat"])  # bit-exact fp32
    assert np.array_equal(np.array([len(c) for c in clusters]), g["cluster_sizes"])
    assert np.array_equal(np.concatenate([c.numpy() for c in clusters]), g["cluster_members"])
    keep = onms.mask_nms(heat > 0.3, 0.6)
    assert np.array_equal(keep.numpy(), g["mask_keep"])
    m = len(boxes)
    ious = onms.iou_aligned(boxes[:m // 2, 1:], boxes[m // 2:2 * (m // 2), 1:])
    assert np.array_equal(ious.numpy(), g["set_ious"])


def test_nms_hand_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_cases.npz"))
    for case in ["identical", "nested_eighth", "disjoint", "zero_volume", "chain"]:
        boxes = torch.from_numpy(g[case + "_boxes"])
        reps, clusters, heat = onms.nms_clustering(boxes, 0.5)
        assert np.array_equal(reps.numpy(), g[case + "_reps"]), case
        assert np.array_equal(heat.numpy(), g[case + "_heat"]), case
    # known answers: nested box has IoU 1/8 (cf. reference utils/metric_util.py:122-126)
    assert abs(float(g["nested_eighth_heat"][0, 1]) - 0.125) < 1e-6
    assert float(g["disjoint_heat"][0, 1]) == 0.0
    assert len(g["identical_reps"]) == 1 and len(g["disjoint_reps"]) == 2
    # chain A~B~C: B joins A's cluster, C survives
    assert list(g["chain_cluster_of"]) == [0, 0, 1]


# ------------------------------------------------------------------------------------------------
# kernel maps
# ------------------------------------------------------------------------------------------------
def _coords(xyz, b=0):
    xyz = np.asarray(xyz, dtype=np.int32)
    return np.concatenate([np.full((len(xyz), 1), b, np.int32), xyz], 1)


def test_kernel_map_single_voxel():
    nbr = so.kernel_map_submanifold(_coords([[5, 5, 5]]), 1, 3)
    assert nbr.shape == (27, 1)
    assert (nbr >= 0).sum() == 1 and nbr[13, 0] == 0


def test_kernel_map_plane_counts():
    xs, ys = np.meshgrid(np.arange(8), np.arange(8), indexing="ij")
    c = _coords(np.stack([xs.ravel(), ys.ravel(), np.zeros(64, int)], 1))
    nbr = so.kernel_map_submanifold(c, 1, 3)
    per_voxel = (nbr >= 0).sum(0)
    interior = (xs.ravel() > 0) & (xs.ravel() < 7) & (ys.ravel() > 0) & (ys.ravel() < 7)
    assert np.all(per_voxel[interior] == 9)          # plane: 9 neighbours per interior voxel
    assert per_voxel.sum() == 6 * 6 * 9 + 4 * 4 + 24 * 6
    # line: 3 per interior voxel
    line = _coords(np.stack([np.arange(10), np.zeros(10, int), np.zeros(10, int)], 1))
    assert (so.kernel_map_submanifold(line, 1, 3) >= 0).sum() == 8 * 3 + 2 * 2


def test_kernel_map_block_2x2x2_stride2():
    g = np.stack(np.meshgrid(np.arange(2), np.arange(2), np.arange(2), indexing="ij"), -1).reshape(-1, 3)
    c = _coords(g + 4)
    coarse, parent = so.downsample_coords(c, 2)
    assert len(coarse) == 1 and np.all(coarse[0] == [0, 4, 4, 4]) and np.all(parent == 0)
    nbr_down, nbr_up = so.kernel_map_stride2(c, parent, 1, 1)
    assert sorted(nbr_down[:, 0].tolist()) == list(range(8))
    for f, (x, y, z) in enumerate(g):
        k = x + 2 * y + 4 * z                        # x fastest
        assert nbr_down[k, 0] == f and nbr_up[k, f] == 0 and (nbr_up[:, f] >= 0).sum() == 1


def test_kernel_map_numpy_equals_dict_random():
    rng = np.random.default_rng(0)
    xyz = np.unique(rng.integers(0, 12, (400, 3)), axis=0)
    c = np.concatenate([_coords(xyz, 0), _coords(xyz[::2], 1)], 0)
    for ts, ks in [(1, 3), (1, 5), (2, 3)]:
        cc = c.copy()
        cc[:, 1:] *= ts
        assert np.array_equal(so.kernel_map_submanifold(cc, ts, ks), so.kernel_map_submanifold_dict(cc, ts, ks))
    # negative coordinates and batch separation
    cn = c.copy()
    cn[:, 1:] -= 6
    a = so.kernel_map_submanifold(cn, 1, 3)
    assert np.array_equal(a, so.kernel_map_submanifold_dict(cn, 1, 3))
    assert np.array_equal(a, so.kernel_map_submanifold(c, 1, 3))      # translation invariance
    # mirrored offsets: nbr[k][o] == i  <=>  nbr[K-1-k][i] == o
    nb = so.kernel_map_submanifold(c, 1, 3)
    for k in range(27):
        o = np.nonzero(nb[k] >= 0)[0]
        assert np.all(nb[26 - k][nb[k][o]] == o)


def test_downsample_sorted_unique_with_negatives():
    rng = np.random.default_rng(1)
    c = _coords(rng.integers(-9, 9, (300, 3)))
    c = np.unique(c, axis=0)
    out, parent = so.downsample_coords(c, 4)
    assert np.all(out[:, 1:] % 4 == 0)
    keys = so.pack_keys(out)
    assert np.all(keys[1:] > keys[:-1])
    assert np.all(out[parent][:, 1:] <= c[:, 1:]) and np.all(c[:, 1:] - out[parent][:, 1:] < 4)


# ------------------------------------------------------------------------------------------------
# sparse conv vs the independent dense oracle (fp64-exact reference, fp32 tolerance stated)
# ------------------------------------------------------------------------------------------------
def _random_sparse(seed, size=14, n=500):
    rng = np.random.default_rng(seed)
    xyz = np.unique(rng.integers(0, size, (n, 3)), axis=0)
    return _coords(xyz)


@pytest.mark.parametrize("ksize", [3, 5])
def test_sparse_conv_matches_dense_stride1(ksize):
    c = _random_sparse(ksize)
    torch.manual_seed(0)
    x = torch.randn(len(c), 4, dtype=torch.float64)
    w = torch.randn(ksize ** 3, 4, 5, dtype=torch.float64)
    y = so.sparse_conv(x, so.kernel_map_submanifold(c, 1, ksize), w)
    ref = so.dense_conv_reference(c, x, w, ksize, 1, False, c)
    assert torch.allclose(y, ref, rtol=1e-12, atol=1e-12)
    y32 = so.sparse_conv(x.float(), so.kernel_map_submanifold(c, 1, ksize), w.float())
    assert torch.allclose(y32.double(), ref, rtol=1e-5, atol=1e-4)     # fp32: rel 1e-5


def test_sparse_conv_matches_dense_stride2_and_transpose():
    c = _random_sparse(7)
    coarse, parent = so.downsample_coords(c, 2)
    nbr_down, nbr_up = so.kernel_map_stride2(c, parent, len(coarse), 1)
    torch.manual_seed(1)
    x = torch.randn(len(c), 3, dtype=torch.float64)
    w = torch.randn(8, 3, 6, dtype=torch.float64)
    y = so.sparse_conv(x, nbr_down, w)
    ref = so.dense_conv_reference(c, x, w, 2, 2, False, coarse)
    assert torch.allclose(y, ref, rtol=1e-12, atol=1e-12)
    # transposed: coarse -> fine
    xc = torch.randn(len(coarse), 6, dtype=torch.float64)
    wt = torch.randn(8, 6, 3, dtype=torch.float64)
    yt = so.sparse_conv(xc, nbr_up, wt)
    reft = so.dense_conv_reference(coarse, xc, wt, 2, 2, True, c)
    assert torch.allclose(yt, reft, rtol=1e-12, atol=1e-12)


def test_dgrad_identity_mirrored_weights():
    """dX of a stride-1 conv == forward conv of dY with W[K-1-k]^T (what the CUDA path packs in mode 1)."""
    c = _random_sparse(3)
    nbr = so.kernel_map_submanifold(c, 1, 3)
    torch.manual_seed(2)
    x = torch.randn(len(c), 4, dtype=torch.float64, requires_grad=True)
    w = torch.randn(27, 4, 5, dtype=torch.float64, requires_grad=True)
    y = so.sparse_conv(x, nbr, w)
    dy = torch.randn_like(y)
    y.backward(dy)
    wd = torch.flip(w.detach(), [0]).transpose(1, 2).contiguous()
    assert torch.allclose(so.sparse_conv(dy, nbr, wd), x.grad, rtol=1e-12, atol=1e-12)
    # wgrad: dW[k] = X[nbr[k]]^T dY
    dw = torch.zeros_like(w)
    for k, (i, o) in enumerate(so.map_to_pairs(nbr)):
        dw[k] = x.detach()[i].t() @ dy[o]
    assert torch.allclose(dw, w.grad, rtol=1e-12, atol=1e-12)


def test_segment_ops():
    torch.manual_seed(0)
    f = torch.randn(50, 4)
    ids = torch.tensor([0] * 1 + [1] * 30 + [2] * 19)
    m = so.segment_mean(f, ids, 3)
    assert torch.allclose(m[0], f[0]) and torch.allclose(m[1], f[1:31].mean(0), atol=1e-6)
    mx = so.segment_max(f, ids, 3)
    assert torch.equal(mx[2], f[31:].max(0)[0])


# ------------------------------------------------------------------------------------------------
# voxelisation restatement (next-row component): the ball-tree result against brute force, invariants
# ------------------------------------------------------------------------------------------------
def test_voxelize_restatement_invariants():
    from oracle import voxelize as ovz
    rng = np.random.default_rng(0)
    p = 3000
    pos = rng.uniform(-0.3, 0.9, (p, 3))
    seg = rng.integers(0, 25, p) * 3
    r = ovz.voxelize_scene(pos, rng.normal(size=(p, 3)).astype(np.float32), rng.normal(size=(p, 3)).astype(np.float32), seg, 0.05)
    v = r["vox_coords"]
    assert np.array_equal(v, np.unique(v, axis=0)) and v.min() >= 0                      # sorted unique, non-negative
    scaled = (pos - min(0, pos.min())) / 0.05
    assert np.array_equal(np.round(scaled), v[r["vox2point"]])                            # vox2point is the inverse map
    d = ((scaled[None, :, :] - v[:, None, :]) ** 2).sum(-1)                               # brute-force nearest point
    assert np.array_equal(d.argmin(1), r["point2vox"])
    assert float(np.sqrt(d.min(1)).max()) <= np.sqrt(3) / 2 + 1e-12                       # within the voxel's own reach
    assert np.array_equal(r["unique_vox_segments"][r["seg2vox"]], r["vox_segments"])
    ids = ovz.to_unique([np.array([5, 5, 9]), np.array([2, 7, 7])])
    assert ids.tolist() == [0, 0, 1, 2, 3, 3]


def test_decode_scene_voxel_semantics_matches_reference(golden_dir):
    """S3DIS branch of detection2mask (per-voxel semantics -> per-segment mode, no mask-NMS): the oracle against the
    fixture produced by the reference's own code (tests/golden/decode_s3dis.npz)."""
    from box2mask_b200.synthetic import label_maps
    from oracle.make_golden import decode_inputs_s3dis, variant_config
    g = np.load(os.path.join(golden_dir, "decode_s3dis.npz"))
    cfg, _, _ = variant_config("s3dis")
    _, _, is_fg = label_maps(13)
    batch, pred = decode_inputs_s3dis()
    boxes = onms.to_boxes(batch["input_location"], pred["mlp_offsets"], pred["mlp_bounds"], torch.sigmoid(pred["mlp_bb_scores"]))
    labels = torch.argmax(pred["mlp_per_vox_semantics"], 1)
    r = onms.decode_scene_voxel_semantics(boxes, labels, batch["vox_segments"][0], batch["seg2vox"][0], is_fg,
                                          cfg.eval_ths[0], cfg.eval_ths[1], cfg.eval_ths[2])
    name = batch["scene"][0]["name"]
    assert np.array_equal(r["conf"].numpy(), g["train_%s_conf" % name])
    assert np.array_equal(r["label_id"], g["train_%s_label_id" % name])
    shape = tuple(g["train_%s_mask_shape" % name])
    ref = np.unpackbits(g["train_%s_mask" % name], axis=1)[:, :shape[1]].astype(bool)
    assert np.array_equal(r["mask"].numpy(), ref)
    assert np.array_equal(r["representatives"].numpy(), g["train_%s_reps" % name])
    ev = np.unpackbits(g["eval_%s_mask" % name], axis=1)[:, :int(g["eval_%s_mask_shape" % name][1])].astype(bool)
    assert np.array_equal(r["mask"][:, batch["vox2point"][0]].numpy(), ev)


def test_label_association_oracle_matches_reference_golden(golden_dir):
    """oracle/label_assoc.py against the outputs of the reference's own ScanNet.approx_association
    (models/dataloader.py:203-314, run by oracle/make_golden.py:golden_label_assoc) for every mode: exact."""
    import numpy as np
    from oracle import label_assoc as la
    g = np.load(os.path.join(golden_dir, "label_assoc.npz"))
    tags = sorted(k[:-4] for k in g.files if k.endswith("_cfg"))
    assert len(tags) == 7
    for tag in tags:
        seed, pa, mv, small, drop, noise = g[tag + "_cfg"]
        labels, scene, unique_segs = la.synthetic_case(int(seed))
        mn, mx, ids, vol = la.prepare_boxes(labels, scene["name"], dropout_boxes=drop, noisy_boxes=noise)
        per_point, per_seg = la.approx_association(scene["positions"], scene["segments"], unique_segs, mn, mx, ids, vol,
                                                   bool(pa), bool(mv), bool(small))
        assert np.array_equal(per_point, g[tag + "_per_point"]), tag
        if per_seg is not None:
            assert np.array_equal(per_seg, g[tag + "_per_seg"]), tag
        else:
            assert tag + "_per_seg" not in g.files
