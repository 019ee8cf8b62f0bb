"""Decode path on the GPU next to the CPU oracle (the reference's torch CPU loops restated):
3-D IoU-NMS clustering with heat-maps on M boxes, heat-map -> bit-packed voxel masks, mask-NMS.
    python tools/decode_bench.py [--m 2000] [--vox 150000]
Results are checked for equality with the oracle while timing."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2mask_b200 import ops  # noqa: E402
from box2mask_b200.decode import NMS_clustering  # noqa: E402
from box2mask_b200.synthetic import make_boxes  # noqa: E402
from oracle import nms as onms  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=2000)
ap.add_argument("--centres", type=int, default=150)
ap.add_argument("--vox", type=int, default=150000)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = "cuda"
boxes = make_boxes(args.m, args.centres, seed=1)
bd = boxes.to(dev)


def gpu_time(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters, out


ms_nms, (reps, clusters, heat) = gpu_time(lambda: NMS_clustering(bd, 0.5))
t0 = time.perf_counter()
r2, c2, h2 = onms.nms_clustering(boxes, 0.5)
cpu_nms = (time.perf_counter() - t0) * 1e3
assert torch.equal(reps.cpu(), r2) and torch.equal(heat.cpu(), h2)
k = len(reps)
print("NMS_clustering  M=%d -> K=%d clusters: GPU %.3f ms (incl. host sync), CPU oracle %.1f ms, bit-exact" % (args.m, k, ms_nms, cpu_nms))

# heat-maps -> voxel masks: every box is a foreground superpoint; voxels map to superpoints at random
rng = np.random.default_rng(0)
seg2vox = torch.from_numpy(rng.integers(0, args.m, args.vox)).to(dev)
fg_rank = torch.arange(args.m, dtype=torch.int32, device=dev)
ms_proj, masks = gpu_time(lambda: ops.heatmap_project(heat, fg_rank, seg2vox, 0.3))
t0 = time.perf_counter()
cpu_masks = (h2[:, seg2vox.cpu()] > 0.3)
cpu_proj = (time.perf_counter() - t0) * 1e3
assert torch.equal(ops.unpack_masks(masks, args.vox).cpu(), cpu_masks)
print("heatmap_project K=%d x N=%d: GPU %.3f ms (%.0f GB/s of the %d MB fp32 mask matrix it replaces), CPU %.1f ms" % (
    k, args.vox, ms_proj, 4.0 * k * args.vox / ms_proj / 1e6, 4 * k * args.vox // 1000000, cpu_proj))
ms_mnms, keep = gpu_time(lambda: ops.mask_nms(masks, 0.6))
t0 = time.perf_counter()
keep_cpu = onms.mask_nms(cpu_masks, 0.6)
cpu_mnms = (time.perf_counter() - t0) * 1e3
assert torch.equal(torch.nonzero(keep).flatten().cpu(), keep_cpu)
print("mask_NMS K=%d x N=%d: GPU %.3f ms, CPU oracle %.1f ms, identical keep list (%d kept)" % (k, args.vox, ms_mnms, cpu_mnms, int(keep.sum())))
