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


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not mounted at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    golden_nms()
