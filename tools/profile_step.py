"""One training step of the bench workload between cudaProfilerStart/Stop (for ncu --profile-from-start off),
plus a per-launch CUDA-event dump of every C-ABI call (kind, shape tag, ms, algorithmic TFLOP/s or GB/s).

    python tools/profile_step.py [--scenes 8] [--scale 0.84] [--dump gpurun_out/launches.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from box2mask_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=8)
ap.add_argument("--scale", type=float, default=0.84)
ap.add_argument("--dump", default=None)
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--cprofile", type=int, default=0, help="cProfile this many steps (host side) and print the top entries")
args = ap.parse_args()

dev = "cuda:0"
scenes = bench.make_scenes(args.scenes, seed=10, scale=args.scale)
rng = np.random.default_rng(0)
model, opt, cfg = bench.build_model(dev, multigpu=False)
batch = bench.jitter_batch(scenes, rng)
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
for _ in range(args.warm):
    bench.train_step(model, opt, batch)
torch.cuda.synchronize()
import time  # noqa: E402
for _ in range(3):   # host enqueue time vs end-to-end time of a step (how far the host runs ahead of the GPU)
    t0 = time.perf_counter()
    bench.train_step(model, opt, batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("step: host returned after %.1f ms, GPU done after %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3))
if args.cprofile:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(args.cprofile):
        bench.train_step(model, opt, batch)
    pr.disable()
    torch.cuda.synchronize()
    for key in ("tottime", "cumulative"):
        print("---- host profile of %d steps, sorted by %s ----" % (args.cprofile, key))
        pstats.Stats(pr).sort_stats(key).print_stats(70)
if args.dump:
    ops.Profile.reset()
    ops.Profile.enabled = True
torch.cuda.profiler.start()
(bench.instrumented_step if args.dump else bench.train_step)(model, opt, batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if args.dump:
    ops.Profile.enabled = False
    rows = []
    for rec in ops.Profile.records:
        kind, fl, by, a, b = rec[:5]
        ms = a.elapsed_time(b)
        rows.append({"kind": kind, "tag": rec[5] if len(rec) > 5 else "", "ms": ms,
                     "tflops": fl / (ms * 1e-3) / 1e12 if fl and ms else None,
                     "gbs": by / (ms * 1e-3) / 1e9 if by and ms else None, "gflop": fl / 1e9})
    json.dump(rows, open(args.dump, "w"))
    rows.sort(key=lambda r: -r["ms"])
    print("total ms in C-ABI calls: %.2f over %d calls" % (sum(r["ms"] for r in rows), len(rows)))
    for r in rows[:45]:
        print("%-22s %-44s %8.3f ms  %s" % (r["kind"], r["tag"], r["ms"],
              ("%.1f TF/s" % r["tflops"]) if r["tflops"] else (("%.0f GB/s" % r["gbs"]) if r["gbs"] else "")))

# ---- phase timeline of one step (CUDA events on the main stream): where the 34 ms go ----
import box2mask_b200 as _pkg  # noqa: E402
ME = _pkg.install_as_minkowski_engine()


def phases():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    names = ["zero_grad", "coordinate maps + weight packing", "trunk + heads forward", "losses", "backward", "optimizer"]
    net = model.net
    torch.cuda.synchronize()
    ev[0].record()
    opt.zero_grad(set_to_none=True)
    ev[1].record()
    sin = ME.SparseTensor(batch["vox_features"], batch["vox_coords"], device=dev)
    ME.prepack_conv_weights(net)
    sin.coordinate_manager.prepare(*net.coordinate_plan())
    ev[2].record()
    b2 = dict(batch)
    b2["_coordinate_manager"] = sin.coordinate_manager
    # compute_loss_detection = forward + losses; split it with an event recorded by a forward hook on the network
    h = net.register_forward_hook(lambda *_: ev[3].record())
    losses, _ = model.compute_loss_detection(b2, epoch=0)
    h.remove()
    ev[4].record()
    losses["optimization_loss"].backward()
    ev[5].record()
    opt.step()
    ev[6].record()
    torch.cuda.synchronize()
    return names, [ev[i].elapsed_time(ev[i + 1]) for i in range(6)]


for _ in range(2):
    names, ms = phases()
print("phase timeline of one training step (ms, CUDA events on the main stream; total %.2f):" % sum(ms))
for nm, t in zip(names, ms):
    print("  %-36s %7.2f" % (nm, t))
