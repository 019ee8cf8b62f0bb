"""Generate tests/golden/*.npz by running the REFERENCE's own Python in the build container.

    python -m oracle.make_golden            # needs /root/reference (not present on the GPU box)

nms_*.npz   : /root/reference/models/iou_nms.py imported unmodified (NMS_clustering, mask_NMS, set_IOUs).
The generating inputs come from box2mask_b200.synthetic (seeded). Fixtures are small and committed.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_ref_module(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_nms():
    from box2mask_b200.synthetic import make_boxes
    ref = load_ref_module("models/iou_nms.py", "ref_iou_nms")
    for name, m, centres, seed, th in [("nms_small", 64, 8, 1, 0.5), ("nms_medium", 400, 30, 2, 0.5),
                                       ("nms_lowth", 300, 20, 3, 0.25)]:
        boxes = make_boxes(m, centres, seed)
        reps, clusters, heat = ref.NMS_clustering(boxes.clone(), cluster_th=th)
        cluster_of = np.full(m, -1, dtype=np.int32)
        for c, members in enumerate(clusters):
            cluster_of[members.numpy()] = c
        # mask NMS on thresholded heat-maps (detection_net.py:446-448), sorted by score as the caller has them
        masks = heat > 0.3
        keep, _ = ref.mask_NMS(masks, 0.6)
        ious = ref.set_IOUs(boxes[:m // 2, 1:], boxes[m // 2:2 * (m // 2), 1:])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), boxes=boxes.numpy(), th=np.float32(th),
                            reps=reps.numpy(), cluster_of=cluster_of, heat=heat.numpy(),
                            cluster_sizes=np.array([len(c) for c in clusters]),
                            cluster_members=np.concatenate([c.numpy() for c in clusters]),
                            mask_keep=keep.numpy(), set_ious=ious.numpy())
        print(name, "clusters", len(reps), "kept masks", len(keep))
    # hand-checkable cases
    cases = {
        "identical": torch.tensor([[0.9, 0, 0, 0, 1, 1, 1], [0.8, 0, 0, 0, 1, 1, 1]], dtype=torch.float32),
        "nested_eighth": torch.tensor([[0.9, 0, 0, 0, 2, 2, 2], [0.8, 0, 0, 0, 1, 1, 1]], dtype=torch.float32),
        "disjoint": torch.tensor([[0.9, 0, 0, 0, 1, 1, 1], [0.8, 2, 2, 2, 3, 3, 3]], dtype=torch.float32),
        "zero_volume": torch.tensor([[0.9, 0, 0, 0, 0, 1, 1], [0.8, 0, 0, 0, 1, 1, 1]], dtype=torch.float32),
        "chain": torch.tensor([[0.9, 0, 0, 0, 1, 1, 1], [0.8, 0.3, 0, 0, 1.3, 1, 1], [0.7, 0.6, 0, 0, 1.6, 1, 1]],
                              dtype=torch.float32),
    }
    out = {}
    for k, b in cases.items():
        reps, clusters, heat = ref.NMS_clustering(b.clone(), cluster_th=0.5)
        out[k + "_boxes"] = b.numpy()
        out[k + "_reps"] = reps.numpy()
        out[k + "_heat"] = heat.numpy()
        cof = np.full(len(b), -1, dtype=np.int32)
        for c, members in enumerate(clusters):
            cof[members.numpy()] = c
        out[k + "_cluster_of"] = cof
    np.savez_compressed(os.path.join(OUT, "nms_cases.npz"), **out)




def golden_net():
    """Run the reference's own SelectionNet / Model code (over oracle/me_shim.py) on a small seeded batch."""
    from box2mask_b200.selection_net import default_config
    from box2mask_b200.synthetic import label_maps, make_batch
    from oracle import me_shim
    from oracle.selection_net import seeded_state_dict
    me_shim.install()
    sys.path.insert(0, REF)
    # the reference hard-codes .to('cuda') in its loss (models/model.py:197,213); run it on the CPU here
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **k)
    torch.Tensor.to = to_cpu
    import models.model as ref_model

    cfg = default_config(mlp_bb_scores_start_epoch=0)
    valid, id2idx, is_fg = label_maps(20)
    model = ref_model.Model(cfg, valid, id2idx, None, is_fg, device="cpu")
    net = model.detection_model
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(seeded_state_dict(shapes, seed=0))
    batch = make_batch(2, seed=7, scale=0.14, density=1.2e4)
    print("golden net batch: voxels", batch["vox_coords"].shape[0], "superpoints", batch["input_location"].shape[0])
    out = {"vox_coords": batch["vox_coords"].numpy(), "vox_features": batch["vox_features"].numpy(),
           "pooling_ids": batch["pooling_ids"].numpy(), "input_location": batch["input_location"].numpy(),
           "gt_bb_offsets": batch["gt_bb_offsets"].numpy(), "gt_bb_bounds": batch["gt_bb_bounds"].numpy(),
           "gt_semantics": batch["gt_semantics"].numpy(), "fg_instances": batch["fg_instances"].numpy(),
           "keys": np.array(list(shapes.keys())), "shapes": np.array([str(s) for s in shapes.values()])}
    # eval-mode forward (running statistics)
    model.eval()
    pred = model.get_prediction(batch, with_grad=False, to_cpu=True, min_size=False)
    for k, v in pred.items():
        out["eval_" + k] = v.numpy()
    # train-mode forward + losses + a few gradients
    model.train()
    losses, pred = model.compute_loss_detection(batch, epoch=0)
    losses["optimization_loss"].backward()
    for k, v in pred.items():
        out["train_" + k] = v.detach().numpy()
    for k in ("optimization_loss", "offset_loss", "bounds_loss", "bb_score_loss", "bb_target_scores"):
        out["loss_" + k] = np.float64(float(losses[k]))
    out["loss_semantics_loss"] = np.float64(float(losses["semantics_loss"]))
    params = dict(net.named_parameters())
    for k in ("conv0p1s1.kernel", "block1.0.conv1.kernel", "block8.1.conv2.kernel", "added_block3.1.conv2.kernel",
              "convtr7p2s2.kernel", "conv2p2s2.kernel", "block2.0.downsample.0.kernel", "bn0.bn.weight",
              "block8.1.norm2.bn.bias", "mlp_offsets.6.kernel", "mlp_semantics.0.bias"):
        g = params[k].grad
        out["grad_" + k] = g.numpy() if g.numel() <= 40000 else g.flatten()[:40000].numpy()
        out["gradnorm_" + k] = np.float64(float(g.norm()))
    torch.Tensor.to = orig_to
    np.savez_compressed(os.path.join(OUT, "selection_net_small.npz"), **out)
    print({k: float(v) for k, v in out.items() if k.startswith("loss_")})

VARIANTS = ("s3dis", "arkit", "maxpool")


def variant_config(kind):
    """(cfg, number of classes, make_batch kwargs) of the extra parity configurations:
    s3dis   : configs/s3dis_fold1.txt (per-voxel semantics head on the un-pooled tensor, 13 classes, score weight 3) plus
              the optional IoU loss of models/model.py:91-129 switched on (use_bb_iou_loss, weight 1), so that L4 is pinned;
    arkit   : configs/arkitscenes.txt (4 cm voxels, 28 classes, loss weights 0.5 / 3 / 0.3);
    maxpool : configs/scannet.txt with max_pool_segments_detection_net (MinkowskiGlobalMaxPooling, detection_net.py:352)."""
    from box2mask_b200.selection_net import default_config
    if kind == "s3dis":
        cfg = default_config(network_heads=["mlp_offsets", "mlp_bounds", "mlp_bb_scores", "mlp_per_vox_semantics"],
                             eval_ths=[0.5, 0.03, 0.3, 0.6], batch_size=4, loss_weight_bb_scores=3,
                             mlp_bb_scores_start_epoch=0, use_bb_iou_loss=True, loss_weight_bb_iou=1.0)
        return cfg, 13, dict(n=2, seed=21, scale=0.14, density=1.2e4, n_classes=13)
    if kind == "arkit":
        cfg = default_config(voxel_size=0.04, eval_ths=[0.5, 0.05, 0.4, 0.6], batch_size=4, loss_weight_bb_scores=3,
                             loss_weight_semantics=0.3, mlp_bb_scores_start_epoch=0)
        return cfg, 28, dict(n=2, seed=22, scale=0.3, density=1.2e4, voxel_size=0.04, n_classes=28)
    if kind == "maxpool":
        cfg = default_config(max_pool_segments_detection_net=True, mlp_bb_scores_start_epoch=0)
        return cfg, 20, dict(n=2, seed=23, scale=0.14, density=1.2e4)
    raise ValueError(kind)


def variant_batch(kind):
    from box2mask_b200.synthetic import make_batch
    cfg, n_cls, kw = variant_config(kind)
    kw = dict(kw)
    batch = make_batch(kw.pop("n"), **kw)
    if kind == "s3dis":     # per-voxel labels = the label of the voxel's superpoint (models/dataloader.py S3DIS loader)
        batch["gt_per_vox_semantics"] = batch["gt_semantics"][batch["pooling_ids"]]
    return cfg, n_cls, batch


def golden_variants():
    """The reference's own Model / SelectionNet (models/model.py, models/detection_net.py, unmodified, over
    oracle/me_shim.py) on the three extra configurations: eval-mode head outputs, train-mode losses (incl. the IoU
    loss and the per-voxel semantics loss) and a few gradients -> tests/golden/selection_net_variants.npz."""
    from box2mask_b200.synthetic import label_maps
    from oracle import me_shim
    from oracle.selection_net import seeded_state_dict
    me_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **k)
    torch.Tensor.to = to_cpu
    import models.model as ref_model
    out = {}
    try:
        for kind in VARIANTS:
            cfg, n_cls, batch = variant_batch(kind)
            valid, id2idx, is_fg = label_maps(n_cls)
            model = ref_model.Model(cfg, valid, id2idx, None, is_fg, device="cpu")
            net = model.detection_model
            shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
            net.load_state_dict(seeded_state_dict(shapes, seed=5))
            out[kind + "_n_keys"] = np.int64(len(shapes))
            model.eval()
            pred = model.get_prediction(batch, with_grad=False, to_cpu=True, min_size=False)
            for k, v in pred.items():
                if k == "vox_feats":
                    continue                     # the un-pooled trunk output (8 MB): the per-voxel head covers it
                # per-voxel outputs: every 4th row keeps the fixture small
                out["%s_eval_%s" % (kind, k)] = v.numpy()[::4] if v.shape[0] > 5000 else v.numpy()
            model.train()
            if kind == "s3dis":
                # IoU loss (models/model.py:91-129): random ground truth never overlaps the predictions (IoU == 0, zero
                # gradient), so the ground-truth boxes of this variant are the train-mode predictions plus noise.
                with torch.no_grad():
                    _, p0 = model.compute_loss_detection(batch, epoch=0)
                g = torch.Generator().manual_seed(99)
                batch["gt_bb_offsets"] = (p0["mlp_offsets"] + 0.05 * torch.randn(p0["mlp_offsets"].shape, generator=g)).detach()
                batch["gt_bb_bounds"] = (p0["mlp_bounds"].clamp(min=0.06) *
                                         (0.8 + 0.4 * torch.rand(p0["mlp_bounds"].shape, generator=g))).detach()
                out["s3dis_gt_bb_offsets"] = batch["gt_bb_offsets"].numpy()
                out["s3dis_gt_bb_bounds"] = batch["gt_bb_bounds"].numpy()
                net.load_state_dict(seeded_state_dict(shapes, seed=5))      # undo the running-statistics update
            losses, pred = model.compute_loss_detection(batch, epoch=0)
            losses["optimization_loss"].backward()
            for k, v in losses.items():
                if k.endswith("_loss") or k == "bb_target_scores":
                    out["%s_loss_%s" % (kind, k)] = np.float64(float(v))
            params = dict(net.named_parameters())
            for k in ("block8.1.conv2.kernel", "mlp_offsets.6.kernel", "mlp_bounds.6.kernel"):
                out["%s_gradnorm_%s" % (kind, k)] = np.float64(float(params[k].grad.norm()))
            print(kind, "voxels", batch["vox_coords"].shape[0], "superpoints", batch["input_location"].shape[0],
                  {k: round(float(v), 5) for k, v in out.items() if k.startswith(kind + "_loss_")})
    finally:
        torch.Tensor.to = orig_to
    np.savez_compressed(os.path.join(OUT, "selection_net_variants.npz"), **out)


def decode_inputs_s3dis(seed=13):
    """One scene (the reference's S3DIS branch only works at batch size 1, detection_net.py:395-396) with per-voxel
    semantics logits, per-scene `vox_segments` and a `vox2point` map."""
    from box2mask_b200.synthetic import make_batch
    batch = make_batch(1, seed=seed, scale=0.3, density=6.0e3, n_classes=13)
    rng = np.random.default_rng(seed)
    loc = batch["input_location"].numpy()
    s = len(loc)
    centres = rng.uniform(0.2, 1.2, (7, 3)).astype(np.float32)
    sizes = rng.uniform(0.15, 0.5, (7, 3)).astype(np.float32)
    which = rng.integers(0, 7, s)
    c = centres[which] + rng.normal(0, 0.03, (s, 3)).astype(np.float32)
    n_vox = batch["vox_coords"].shape[0]
    seg = batch["pooling_ids"].numpy()
    # per-voxel logits: a per-segment preferred class plus noise, so that the per-segment mode is non-trivial
    seg_cls = rng.integers(0, 13, s)
    logits = rng.normal(0, 1, (n_vox, 13)).astype(np.float32)
    logits[np.arange(n_vox), seg_cls[seg]] += 1.5
    pred = {
        "mlp_offsets": torch.from_numpy((c - loc).astype(np.float32)),
        "mlp_bounds": torch.from_numpy((sizes[which] * rng.uniform(0.85, 1.15, (s, 3))).astype(np.float32)),
        "mlp_bb_scores": torch.from_numpy(rng.normal(0, 2, (s, 1)).astype(np.float32)),
        "mlp_per_vox_semantics": torch.from_numpy(logits),
    }
    batch["vox_segments"] = [seg.copy()]
    batch["vox2point"] = [torch.from_numpy(rng.integers(0, n_vox, int(1.5 * n_vox))).long()]
    return batch, pred


def golden_decode_s3dis():
    """The reference's detection2mask with requires_voxel_outputs=True (per-voxel semantics -> per-segment torch.mode,
    no mask-NMS; models/detection_net.py:378-381,395-415,449-451) -> tests/golden/decode_s3dis.npz."""
    from types import SimpleNamespace
    from box2mask_b200.synthetic import label_maps
    from oracle import me_shim
    me_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.detection_net as ref_det
    cfg, _, _ = variant_config("s3dis")
    valid, _, is_fg = label_maps(13)
    batch, pred = decode_inputs_s3dis()
    stub = SimpleNamespace(semantic_valid_class_ids=valid, is_foreground=is_fg, requires_voxel_outputs=True)
    if not hasattr(np, "bool"):          # the reference uses the removed alias np.bool (detection_net.py:451)
        np.bool = bool
    out = {}
    for mode in ("eval", "train"):
        res = ref_det.SelectionNet.detection2mask(stub, batch, {k: v.clone() for k, v in pred.items()}, cfg, mode, True,
                                                  *cfg.eval_ths)
        for name, r in res.items():
            out["%s_%s_conf" % (mode, name)] = np.asarray(r["conf"], dtype=np.float32)
            out["%s_%s_label_id" % (mode, name)] = np.asarray(r["label_id"], dtype=np.int32)
            out["%s_%s_mask" % (mode, name)] = np.packbits(np.asarray(r["mask"], dtype=bool), axis=1)
            out["%s_%s_mask_shape" % (mode, name)] = np.array(r["mask"].shape)
            if mode == "train":
                out["train_%s_reps" % name] = np.asarray(r["cluster_representatives"], dtype=np.int64)
            print("s3dis", mode, name, "instances", len(r["conf"]), "mask", tuple(r["mask"].shape))
    np.savez_compressed(os.path.join(OUT, "decode_s3dis.npz"), **out)


def decode_inputs(seed=11):
    """A small seeded batch + head outputs whose boxes form clusters (shared by the golden and the tests)."""
    from box2mask_b200.synthetic import make_batch
    batch = make_batch(2, seed=seed, scale=0.3, density=6.0e3)
    rng = np.random.default_rng(seed)
    loc = batch["input_location"].numpy()
    s = len(loc)
    centres = rng.uniform(0.2, 1.2, (2, 7, 3)).astype(np.float32)          # 7 object centres per scene
    sizes = rng.uniform(0.15, 0.5, (2, 7, 3)).astype(np.float32)
    which = rng.integers(0, 7, s)
    b = batch["batch_ids"].numpy()
    c = centres[b, which] + rng.normal(0, 0.03, (s, 3)).astype(np.float32)
    pred = {
        "mlp_offsets": torch.from_numpy((c - loc).astype(np.float32)),
        "mlp_bounds": torch.from_numpy((sizes[b, which] * rng.uniform(0.85, 1.15, (s, 3))).astype(np.float32)),
        "mlp_bb_scores": torch.from_numpy(rng.normal(0, 2, (s, 1)).astype(np.float32)),
        "mlp_semantics": torch.from_numpy(rng.normal(0, 1, (s, 20)).astype(np.float32)),
    }
    batch["vox2point"] = [torch.from_numpy(rng.integers(0, len(sv), int(1.5 * len(sv)))).long() for sv in batch["seg2vox"]]
    return batch, pred


def golden_decode():
    """The reference's own detection2mask (models/detection_net.py:369-488, CPU torch loops) on decode_inputs()."""
    from types import SimpleNamespace
    from box2mask_b200.selection_net import default_config
    from box2mask_b200.synthetic import label_maps
    from oracle import me_shim
    me_shim.install()
    sys.path.insert(0, REF)
    import models.detection_net as ref_det
    cfg = default_config()
    valid, _, is_fg = label_maps(20)
    batch, pred = decode_inputs()
    stub = SimpleNamespace(semantic_valid_class_ids=valid, is_foreground=is_fg, requires_voxel_outputs=False)
    out = {}
    for mode in ("eval", "train"):
        res = ref_det.SelectionNet.detection2mask(stub, batch, {k: v.clone() for k, v in pred.items()}, cfg, mode, True,
                                                  *cfg.eval_ths)
        for name, r in res.items():
            out["%s_%s_conf" % (mode, name)] = np.asarray(r["conf"], dtype=np.float32)
            out["%s_%s_label_id" % (mode, name)] = np.asarray(r["label_id"], dtype=np.int32)
            out["%s_%s_mask" % (mode, name)] = np.packbits(np.asarray(r["mask"], dtype=bool), axis=1)
            out["%s_%s_mask_shape" % (mode, name)] = np.array(r["mask"].shape)
            if mode == "train":
                out["train_%s_reps" % name] = np.asarray(r["cluster_representatives"], dtype=np.int64)
                out["train_%s_bbs" % name] = np.asarray(r["bbs"], dtype=np.float32)
            print(mode, name, "instances", len(r["conf"]), "mask", tuple(r["mask"].shape))
    np.savez_compressed(os.path.join(OUT, "decode_small.npz"), **out)


def golden_label_assoc():
    """tests/golden/label_assoc.npz: the reference's own `ScanNet.approx_association` (models/dataloader.py:203-314), called
    unbound on seeded synthetic scenes for every mode. The module's other imports (open3d, pyviz3d, the dataset readers,
    MinkowskiEngine) are not needed by that method and are stubbed for the import; `np.int` (removed from numpy) and the
    old scipy.stats.mode return shape are shimmed, the method itself runs unmodified."""
    import types
    import scipy.stats as stats
    from oracle import me_shim
    from oracle.label_assoc import synthetic_case
    me_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ("dataprocessing.scannet", "dataprocessing.arkitscenes", "dataprocessing.s3dis"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "dataprocessing" not in sys.modules:
        pkg = types.ModuleType("dataprocessing")
        pkg.__path__ = []
        sys.modules["dataprocessing"] = pkg
    if not hasattr(np, "int"):
        np.int = int
    if "numpy.lib.type_check" not in sys.modules:       # an unused private import at the top of the reference module
        tc = types.ModuleType("numpy.lib.type_check")
        tc._is_type_dispatcher = None
        sys.modules["numpy.lib.type_check"] = tc
    orig_mode = stats.mode

    def old_mode(a, axis=0, **kw):          # scipy < 1.9 returned 1-element arrays for axis=None
        r = orig_mode(np.asarray(a), axis=axis, keepdims=False)
        return np.atleast_1d(r.mode), np.atleast_1d(r.count)
    stats.mode = old_mode
    try:
        import models.dataloader as ref_dl
        out = {}
        cases = [("seg", False, False, False, 0.0, 0.0), ("seg_small", False, False, True, 0.0, 0.0),
                 ("point", True, False, False, 0.0, 0.0), ("point_small", True, False, True, 0.0, 0.0),
                 ("vote", False, True, False, 0.0, 0.0), ("vote_small", False, True, True, 0.0, 0.0),
                 ("seg_noise_drop", False, False, True, 0.3, 0.1)]
        for ci, (tag, pa, mv, small, drop, noise) in enumerate(cases):
            labels, scene, unique_segs = synthetic_case(seed=40 + ci)
            cfg = types.SimpleNamespace(dropout_boxes=drop, noisy_boxes=noise, smallest_bb_heuristic=small)
            fake_self = types.SimpleNamespace(cfg=cfg)
            per_point, per_seg = ref_dl.ScanNet.approx_association(fake_self, labels, scene, pa, mv, unique_segs, {})
            out[tag + "_per_point"] = np.asarray(per_point).astype(np.int64)
            if per_seg is not None:
                out[tag + "_per_seg"] = np.asarray(per_seg).astype(np.int64)
            out[tag + "_cfg"] = np.array([40 + ci, int(pa), int(mv), int(small), drop, noise], dtype=np.float64)
            print("label association golden", tag, "points", len(per_point), "labels", np.unique(per_point)[:8])
        np.savez_compressed(os.path.join(OUT, "label_assoc.npz"), **out)
    finally:
        stats.mode = orig_mode


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not mounted at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, fn in (("nms", golden_nms), ("net", golden_net), ("decode", golden_decode), ("variants", golden_variants),
                     ("decode_s3dis", golden_decode_s3dis), ("label_assoc", golden_label_assoc)):
        if not only or name in only:
            fn()
