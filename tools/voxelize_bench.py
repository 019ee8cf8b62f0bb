"""Voxelisation + collate of ScanNet-shape scenes: GPU (box2mask_b200.voxelize) against the CPU restatement of the
reference's data loader (oracle/voxelize.py: numpy + scikit-learn ball tree), scenes/s on the box's own cores.
    python tools/voxelize_bench.py [--scenes 8] [--points 240000]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2mask_b200 import voxelize as vz  # noqa: E402
from box2mask_b200.synthetic import make_scene  # noqa: E402
from oracle import voxelize as ovz  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=8)
ap.add_argument("--cpu-scenes", type=int, default=2)
args = ap.parse_args()
clouds = []
for i in range(args.scenes):
    s = make_scene(40000 + i, scale=0.84)
    rng = np.random.default_rng(i)
    c = s["vox_coords"].astype(np.float64) * 0.02
    # ~1.6 points per voxel: the voxel centres jittered, plus a second jittered copy of 60 % of them
    extra = rng.random(len(c)) < 0.6
    pos = np.concatenate([c + rng.uniform(-0.009, 0.009, c.shape), c[extra] + rng.uniform(-0.009, 0.009, (int(extra.sum()), 3))], 0)
    seg = np.concatenate([s["vox_segments"], s["vox_segments"][extra]], 0)
    col = rng.normal(size=(len(pos), 3)).astype(np.float32)
    nor = rng.normal(size=(len(pos), 3)).astype(np.float32)
    clouds.append((pos, col, nor, seg))
print("scenes %d, points/scene %d" % (len(clouds), int(np.mean([len(c[0]) for c in clouds]))))
dev = "cuda"
pinned = [tuple(torch.from_numpy(a).pin_memory() for a in c) for c in clouds]


def gpu_batch():
    items = [vz.voxelize_scene(*(t.to(dev, non_blocking=True) for t in c), 0.02) for c in pinned]
    return vz.collate_scenes(items)


b = gpu_batch(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    b = gpu_batch()
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / 3
print("GPU (H2D of the points included): %.1f ms per batch of %d scenes = %.0f scenes/s; voxels %d" % (
    t_gpu * 1e3, len(clouds), len(clouds) / t_gpu, b["vox_coords"].shape[0]))
t0 = time.perf_counter()
for c in clouds[:args.cpu_scenes]:
    ovz.voxelize_scene(*c, 0.02)
t_cpu = (time.perf_counter() - t0) / args.cpu_scenes
print("CPU restatement (numpy + scikit-learn ball tree, 1 process): %.2f s per scene = %.2f scenes/s" % (t_cpu, 1 / t_cpu))
