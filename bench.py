#!/usr/bin/env python
"""Benchmark of the Box2Mask hot path on B200: SelectionNet ("Res16UNet34C") training step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = the reference training step (models/training.py:63-70) on one batch of 8 synthetic
ScanNet-shape scenes per GPU (configs/scannet.txt: batch_size 8, ~150k voxels at 2 cm each): build the
coordinate hash and all 16 kernel maps, forward, box-vote losses, backward, (gradient all-reduce,) Adam step.
Prints ONE JSON line (contract in the task statement): value = whole-job scenes/s with inputs resident in HBM,
e2e = the same through the public API from pinned host buffers incl. H2D copies and the D2H read of the loss,
roofline = convolution kernels (tcgen05) against the measured bf16 peak, cpu_baseline = the CPU oracle
("ME-CPU-algorithm restatement", MinkowskiEngine 0.5.4 is not installable offline) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENES_PER_GPU = 8
SCENE_SCALE = 0.84          # ~150k voxels at 2 cm
METRIC = "scenes/sec Res16UNet34C fwd+bwd (2cm ScanNet-shape)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "gbs": d["hbm_gbs"], "src": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "gbs": 6650.0, "src": "fallback"}


def conv_traffic():
    """(DRAM bytes (read + write) per convolution launch, source file): averaged over the conv launches of one training
    step, from the latest committed ncu capture by name (profiles/*conv_dram*.json, written by tools/ncu_conv_traffic.py). It is
    NOT measured in this run (ncu cannot run inside a timed benchmark), so the line names the file it came from."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*conv_dram*.json")), key=os.path.basename)   # r1* < r2_* < r2f_*
    if not files:
        return None, None
    try:
        return float(json.load(open(files[-1]))["dram_bytes_per_launch"]), "profiles/" + os.path.basename(files[-1])
    except Exception:
        return None, None


def make_scenes(n, seed, scale):
    from box2mask_b200.synthetic import make_scene
    return [make_scene(1000 * seed + i, scale=scale) for i in range(n)]



def balanced_scenes(n, seed, scale, world, dev):
    """Per-rank scenes. On several GPUs every step waits for the rank with the most voxels (the ranks draw different
    scenes: 144 k - 164 k voxels each), so each rank generates n + 2 candidates and keeps the n whose total voxel count is
    closest to n x the all-rank mean (what a size-bucketing sampler does for a real data set; no scene is altered and
    the single-GPU workload is unchanged)."""
    if world <= 1:
        return make_scenes(n, seed=seed, scale=scale)
    import itertools
    import torch.distributed as dist
    pool = make_scenes(n + 2, seed=seed, scale=scale)
    counts = [len(sc["vox_coords"]) for sc in pool]
    mean = torch.tensor([float(np.mean(counts))], device=dev)
    dist.all_reduce(mean)
    target = n * float(mean.item()) / world
    best = min(itertools.combinations(range(len(pool)), n), key=lambda c: abs(sum(counts[i] for i in c) - target))
    return [pool[i] for i in best]


def jitter_batch(scenes, rng):
    """A fresh batch from cached scenes: random scene order and a random integer translation per scene
    (different coordinates every step, so no coordinate map could be reused)."""
    from box2mask_b200.synthetic import collate
    out = []
    for i in rng.permutation(len(scenes)):
        s = dict(scenes[i])
        s["vox_coords"] = s["vox_coords"] + rng.integers(0, 64, (1, 3)).astype(np.int32)
        out.append(s)
    b = collate(out)
    b["num_segments"] = int(b["input_location"].shape[0])
    return b


TENSOR_KEYS = ("vox_coords", "vox_features", "pooling_ids", "input_location", "gt_bb_offsets", "gt_bb_bounds",
               "gt_semantics", "fg_instances", "fg_index")


class ClockSampler:
    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_model(device, multigpu, grad_sync="flat", trunk_executor=True, overlap_wgrad=True):
    from box2mask_b200.model import Model
    from box2mask_b200.selection_net import default_config
    from box2mask_b200.synthetic import label_maps
    cfg = default_config(multigpu=multigpu, mlp_bb_scores_start_epoch=0, grad_sync=grad_sync, trunk_executor=trunk_executor)
    valid, id2idx, is_fg = label_maps(20)
    torch.manual_seed(0)
    model = Model(cfg, valid, id2idx, None, is_fg, device=device)
    if trunk_executor:
        model.net.trunk_executor().overlap_wgrad = overlap_wgrad
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)     # models/training.py:37-38
    return model, opt, cfg


def train_step(model, opt, batch):
    opt.zero_grad(set_to_none=True)
    losses, _ = model.compute_loss_detection(batch, epoch=0)
    loss = losses["optimization_loss"]
    loss.backward()
    opt.step()
    return loss


def instrumented_step(model, opt, batch):
    """train_step for the per-launch CUDA-event pass. An event pair around a launch measures the kernel only while the
    GPU has a backlog: a deep-level launch runs 10-25 us but takes the host ~50 us to enqueue (more with two events per
    call), so on an idle GPU the pair measures the host. The coordinate maps (whose row counts are read back) are
    therefore built first, and a spin kernel ahead of the forward and of the backward pass lets the host enqueue the whole
    pass before the GPU starts on it; the spins sit outside every event pair."""
    opt.zero_grad(set_to_none=True)
    batch = dict(batch)
    model.prefetch_coordinates(batch)
    torch.cuda.synchronize()
    torch.cuda._sleep(60_000_000)          # ~30 ms at 1.965 GHz
    losses, _ = model.compute_loss_detection(batch, epoch=0)
    loss = losses["optimization_loss"]
    torch.cuda._sleep(120_000_000)
    loss.backward()
    opt.step()
    return loss


def cpu_reference_step(scenes, threads):
    """The CPU oracle (ME-CPU-algorithm restatement): fwd + losses + bwd + Adam step on a bounded sample; returns
    seconds."""
    from box2mask_b200.selection_net import SelectionNet, default_config
    from box2mask_b200.synthetic import collate, label_maps
    from oracle.selection_net import OracleNet, detection_loss, seeded_state_dict
    torch.set_num_threads(threads)
    cfg = default_config(mlp_bb_scores_start_epoch=0)
    net = SelectionNet(cfg, "cpu", list(range(20)), out_channels=[96, 96, 6])
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()})
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    _, id2idx, _ = label_maps(20)
    b = collate(scenes)
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-3)      # models/training.py:37-38
    t0 = time.time()
    out = OracleNet(sd, cfg, training=True).forward(b["vox_coords"].numpy(), b["vox_features"], b["pooling_ids"])
    detection_loss(out, b, cfg, 0, id2idx)["optimization_loss"].backward()
    opt.step()
    return time.time() - t0


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(16, os.cpu_count() or 1)        # config_loader.py:3-4 sets OMP_NUM_THREADS=16 "for the ME engine"
    scenes = make_scenes(2, seed=1, scale=SCENE_SCALE)
    times = []
    for i in range(args.warmup + args.steps):
        t = cpu_reference_step(scenes, threads)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.median(times))
    val = len(scenes) / (ms / 1e3)
    sample = "2 ScanNet-shape scenes (%d voxels) fwd+loss+bwd+Adam per step (median of %d steps after %d warm-up), oracle " \
             "port of ME's CPU algorithm" % (sum(len(s["vox_coords"]) for s in scenes), len(times), args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "scenes/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs/scannet.txt training step (fwd+bwd+box-vote loss), ScanNet-shape scenes, "
                               "bounded sample of 2 scenes/step on CPU (per-scene throughput; the GPU arm steps 8 scenes)"},
        "cpu_baseline": {"value": val, "unit": "scenes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


EVAL_METRIC = "scenes/sec Res16UNet34C eval: forward + box-vote decode + 3D IoU-NMS + superpoint pooling (2cm ScanNet-shape)"


def run_eval(args):
    """--workload eval (BASELINE configs 1 and 3): models/evaluation.py:70-98 - forward in eval mode (BatchNorm folded into
    the convolution epilogues), superpoint pooling, heads, then detection2mask per scene (vote boxes -> 3-D IoU-NMS
    clustering with heat-maps -> voxel masks -> mask-NMS -> label vote). Scenes are sharded over the ranks with no
    communication. One step = one batch of --scenes scenes per GPU."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from box2mask_b200 import ops
    ops._lib.load()
    args.warmup = max(args.warmup, 3)
    scenes = balanced_scenes(args.scenes, 10 + rank, args.scale, world, dev)
    rng = np.random.default_rng(rank)
    model, _, cfg = build_model(dev, multigpu=False)
    model.eval()
    host_batches = [jitter_batch(scenes, rng) for _ in range(min(args.warmup + args.steps, 4))]
    for b in host_batches:
        for k in TENSOR_KEYS:
            b[k] = b[k].pin_memory()
        # Random-initialised heads do not vote coherently (every superpoint would become its own cluster), so the decode
        # stage is fed votes of the shape a trained network produces: 24 synthetic instances per scene, every superpoint
        # votes for the box of the nearest instance centre (+ 2 cm noise), score logits in (-1, 3), the instance's class.
        # The network's own head outputs are still computed inside the timed region.
        g = torch.Generator().manual_seed(0)
        loc, bid = b["input_location"], b["batch_ids"]
        off, bnd = torch.zeros_like(loc), torch.zeros_like(loc)
        sem = torch.zeros(loc.shape[0], dtype=torch.long)
        for sc in range(int(bid.max()) + 1):
            m = bid == sc
            lo, hi = loc[m].min(0)[0], loc[m].max(0)[0]
            ctr = lo + (hi - lo) * torch.rand((24, 3), generator=g)
            half = 0.2 + 0.6 * torch.rand((24, 3), generator=g)
            cls = torch.randint(0, 20, (24,), generator=g)
            near = torch.cdist(loc[m], ctr).argmin(1)
            off[m] = ctr[near] - loc[m]
            bnd[m] = half[near]
            sem[m] = cls[near]
        b["_votes"] = {
            cfg.mlp_offsets: off + 0.02 * torch.randn(off.shape, generator=g),
            cfg.mlp_bounds: (bnd + 0.02 * torch.randn(bnd.shape, generator=g)).clamp(min=0.04),
            cfg.mlp_bb_scores: 4.0 * torch.rand((loc.shape[0], 1), generator=g) - 1.0,
            cfg.mlp_semantics: torch.nn.functional.one_hot(sem, 20).float() * 8.0,
        }
    dev_batches = []
    for b in host_batches:
        d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items() if k != "_votes"}
        d["_votes"] = {k: v.to(dev) for k, v in b["_votes"].items()}
        dev_batches.append(d)
    voxels = float(np.mean([b["vox_coords"].shape[0] for b in host_batches]))
    stage_ms = {"forward": 0.0, "decode": 0.0, "n": 0}

    def step(b, e2e):
        if e2e:
            votes = {k: v.to(dev, non_blocking=True) for k, v in b["_votes"].items()}
            b = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in b.items() if k != "_votes"}
        else:
            votes = b["_votes"]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        pred = model.get_prediction(b, with_grad=False, to_cpu=False)
        e[1].record()
        res = model.pred2mask(b, dict(pred, **votes), "train")          # masks per scene (device work + the result copy)
        e[2].record()
        return res, e

    def timed(batches, steps, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs = []
        e0.record()
        n_inst = 0
        for i in range(steps):
            res, e = step(batches[i % len(batches)], e2e)
            evs.append(e)
            n_inst += sum(len(r["conf"]) for r in res.values())
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for e in evs:
            stage_ms["forward"] += e[0].elapsed_time(e[1]); stage_ms["decode"] += e[1].elapsed_time(e[2]); stage_ms["n"] += 1
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, n_inst / steps

    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(max(args.warmup, len(dev_batches))):
        step(dev_batches[i % len(dev_batches)], False)
    REGIONS = max(1, args.regions)
    ops.Profile.reset()
    runs = [timed(dev_batches, args.steps, False) for _ in range(REGIONS)]
    launches = ops.Profile.launches
    ms_all = [r[0] for r in runs]
    ms_step = float(np.median(ms_all))
    fwd_ms, dec_ms = stage_ms["forward"] / stage_ms["n"], stage_ms["decode"] / stage_ms["n"]
    clocks = sampler.stop() if sampler else None
    ms_e2e_all = [timed(host_batches, args.steps, True)[0] for _ in range(REGIONS)]
    ms_e2e = float(np.median(ms_e2e_all))
    h2d = int(np.mean([sum(b[k].numel() * b[k].element_size() for k in TENSOR_KEYS) +
                       sum(v.numel() * v.element_size() for v in b["_votes"].values()) for b in host_batches]))
    # roofline of the forward convolutions: instrumented pass, CUDA events around every launch
    ops.Profile.reset()
    ops.Profile.enabled = True
    for i in range(2):
        step(dev_batches[i % len(dev_batches)], False)
    torch.cuda.synchronize()
    ops.Profile.enabled = False
    agg = {}
    for kind, fl, by, a, b_, _tag in ops.Profile.records:
        d = agg.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += a.elapsed_time(b_); d["flops"] += fl; d["bytes"] += by; d["launches"] += 1
    pk = peaks()
    conv = agg.get("conv_forward", {"ms": 0.0, "flops": 0.0, "launches": 0})
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] else 0.0
    traffic, traffic_src = conv_traffic()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total = args.scenes * world
    d2h = int(runs[0][1] * voxels / args.scenes) if runs else 0      # bool voxel masks of the instances (+ scores, labels)
    print(json.dumps({
        "metric": EVAL_METRIC, "value": total / (ms_step * 1e-3), "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs/scannet.txt evaluation (BASELINE configs 1/3): %d ScanNet-shape scenes/GPU (%.0f "
                               "voxels/GPU at 2 cm): hash + 16 kernel maps + eval forward (BatchNorm folded) + pooling + heads "
                               "+ detection2mask per scene; no communication" % (args.scenes, voxels),
                   "scenes_per_gpu": args.scenes, "voxels_per_gpu": voxels, "parallelism": "dp%d (scene-sharded)" % world,
                   "stage_ms": {"forward": fwd_ms, "decode": dec_ms},
                   "instances_per_step": runs[0][1],
                   "votes": "decode fed coherent synthetic votes (24 instances/scene, nearest-centre assignment, 2 cm noise; "
                            "random-init heads do not vote coherently); the network's own head outputs are computed inside "
                            "the timed region",
                   "timing": "median of %d regions of %d steps each, ms/step: value %s, e2e %s" % (
                       REGIONS, args.steps, ["%.2f" % m for m in ms_all], ["%.2f" % m for m in ms_e2e_all]),
                   "cache": "inputs larger than L2 (full-resolution activations >= 235 MB), freshly translated coordinates "
                            "every step"},
        "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "conv_fwd_kernel (%d launches/step)" % (conv["launches"] // 2),
                     "kernel_ms_per_step": conv["ms"] / 2, "algorithmic_flops_per_step": conv["flops"] / 2},
        "cpu_baseline": None,
        "kernels": {k: {"ms_per_step": v["ms"] / 2, "launches_per_step": v["launches"] / 2} for k, v in
                    sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
    }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2m")
    ap.add_argument("--scenes", type=int, default=SCENES_PER_GPU)
    ap.add_argument("--scale", type=float, default=SCENE_SCALE)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-bn", action="store_true")
    ap.add_argument("--grad-sync", default="overlap", choices=["overlap", "flat", "ddp"],
                    help="flat: one all-reduce of the flat gradient buffer after backward (box2mask_b200/grad_sync.py); "
                         "ddp: torch DistributedDataParallel buckets overlapped with backward")
    ap.add_argument("--regions", type=int, default=3, help="timed regions of --steps steps each; the median is reported")
    ap.add_argument("--decode-threads", type=int, default=0, help="host threads of the box-vote decode (decode.DECODE_THREADS)")
    ap.add_argument("--flush-every", type=int, default=0, help="commands per b2m_run_commands call (ops.LaunchList.FLUSH_EVERY)")
    ap.add_argument("--hp-stream", action="store_true",
                    help="run the step on a high-priority stream (experiment: with --prefetch the map construction of the "
                         "next step, on a default-priority stream, then only fills SMs the step leaves idle)")
    ap.add_argument("--prefetch", action="store_true",
                    help="build the coordinate maps one step ahead on a side stream (Model.prefetch_coordinates); measured "
                         "slower than building them inside the step: the persistent conv kernels leave the side stream no SMs")
    ap.add_argument("--workload", default="train", choices=["train", "eval"],
                    help="train: configs/scannet.txt training step (the headline metric); eval: forward + decode")
    ap.add_argument("--no-trunk-executor", action="store_true", help="run the trunk module by module through autograd")
    ap.add_argument("--no-wgrad-overlap", action="store_true", help="weight gradients on the main stream")
    ap.add_argument("--gather-mode", default="cpasync", choices=["cpasync", "tma"],
                    help="how the convolution kernels fetch feature rows (b2m_set_option B2M_OPT_GATHER_MODE)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.decode_threads:
        from box2mask_b200 import decode as _decode
        _decode.DECODE_THREADS = args.decode_threads
    if args.workload == "eval":
        return run_eval(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and args.grad_sync == "ddp":
        # leave SMs for the NCCL all-reduce kernels that overlap the backward pass (persistent conv kernels otherwise
        # hold every SM and the collective forces a second wave of their CTAs)
        from box2mask_b200 import _lib as _b2m_lib
        _b2m_lib.set_option(_b2m_lib.OPT_MAX_CTAS, 132)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from box2mask_b200 import ops
    ops._lib.load()      # fail loudly if the CUDA library is missing
    ops._lib.set_option(ops._lib.OPT_GATHER_MODE, 1 if args.gather_mode == "tma" else 0)

    scenes = balanced_scenes(args.scenes, 10 + rank, args.scale, world, dev)
    rng = np.random.default_rng(rank)
    if args.flush_every:
        ops.LaunchList.FLUSH_EVERY = args.flush_every

    if args.hp_stream:
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
    model, opt, cfg = build_model(dev, multigpu=world > 1, grad_sync=args.grad_sync,
                                  trunk_executor=not args.no_trunk_executor, overlap_wgrad=not args.no_wgrad_overlap)
    if world > 1 and not args.sync_bn:
        for m in model.net.modules():      # per-rank BatchNorm statistics unless --sync-bn
            if hasattr(m, "process_group"):
                m.process_group = None
    n_total = args.warmup + args.steps
    host_batches = [jitter_batch(scenes, rng) for _ in range(min(n_total, 4))]
    for b in host_batches:
        for k in TENSOR_KEYS:
            b[k] = b[k].pin_memory()
    dev_batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()} for b in host_batches]
    voxels = float(np.mean([b["vox_coords"].shape[0] for b in host_batches]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, read_loss):
        # Steady-state pipeline: while the GPU works on step i the host builds the coordinate maps of step i+1 on a
        # side stream (Model.prefetch_coordinates). The maps of the first timed step are built before the clock
        # starts and the last timed iteration builds those of the step after the region, so the region holds exactly
        # `steps` steps and `steps` map constructions.
        if args.prefetch:
            model.prefetch_coordinates(batches[0])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        staged = model.stage_batch(batches[0]) if read_loss else None      # (inside the timed region, like every copy)
        for i in range(steps):
            src = batches[i % len(batches)]
            cm = src.pop("_coordinate_manager", None)
            b = src
            if read_loss:
                # end to end: H2D of this step's inputs from pinned memory, D2H of the loss. The copy of step i + 1 is
                # issued on a copy stream right after step i has been enqueued (Model.stage_batch - what a prefetching
                # data loader does), so all `steps` transfers lie inside the region and all but the first run under
                # the previous step's kernels.
                b = staged
            if cm is not None:
                b = dict(b)
                b["_coordinate_manager"] = cm
            loss = train_step(model, opt, b)
            if read_loss and i + 1 < steps:
                staged = model.stage_batch(batches[(i + 1) % len(batches)])
            if args.prefetch:
                model.prefetch_coordinates(batches[(i + 1) % len(batches)])
            if read_loss:
                float(loss.item())
        e1.record()
        batches[steps % len(batches)].pop("_coordinate_manager", None)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # warm-up: at least W steps and every distinct batch once (their row counts differ, so each one grows the
    # caching allocator the first time it is seen; that must not happen inside the timed region)
    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML initialisation) takes driver locks for a few hundred
    # milliseconds, which once landed inside the first timed region and stalled kernel submission (72 instead of 44 ms
    # per step); by the time the warm-up is over it only polls.
    sampler = ClockSampler(local) if rank == 0 else None
    # Untimed warm-up: the W steps asked for, and at least three passes over the distinct batches - the caching allocator
    # keeps growing for a few steps (weight gradients run on a side stream, so blocks are handed back late), and a
    # region that still contains that growth measured 37-54 ms/step next to 33-34 ms ones.
    n_warm = max(args.warmup, 3 * len(dev_batches))
    for i in range(n_warm):
        train_step(model, opt, dev_batches[i % len(dev_batches)])
    for i in range(len(host_batches)):          # the end-to-end path allocates its own device copies: warm that too
        hb = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in host_batches[i].items()}
        float(train_step(model, opt, hb).item())
    args.warmup = n_warm + len(host_batches)    # reported as done
    # Each region times exactly `steps` steps between barrier + synchronize; the MEDIAN of REGIONS regions is reported
    # and every region's figure is kept in config.timing.
    REGIONS = max(1, args.regions)
    ms_all = []
    for r in range(REGIONS):
        ops.Profile.reset()
        ms_all.append(timed(dev_batches, args.steps, read_loss=False))
    launches = ops.Profile.launches
    ms_step = float(np.median(ms_all))
    clocks = sampler.stop() if sampler else None
    ms_e2e_all = [timed(host_batches, args.steps, read_loss=True) for _ in range(REGIONS)]
    ms_e2e = float(np.median(ms_e2e_all))
    h2d = int(np.mean([sum(b[k].numel() * b[k].element_size() for k in TENSOR_KEYS) for b in host_batches]))

    # instrumented pass: CUDA events around every C-ABI launch (same stream), for the roofline figures
    ops.Profile.reset()
    ops.Profile.enabled = True
    prof_steps = min(args.steps, 2)
    for i in range(prof_steps):
        instrumented_step(model, opt, dev_batches[i % len(dev_batches)])
    torch.cuda.synchronize()
    ops.Profile.enabled = False
    agg = {}
    for kind, fl, by, a, b_, _tag in ops.Profile.records:
        d = agg.setdefault(kind, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        d["ms"] += a.elapsed_time(b_); d["flops"] += fl; d["bytes"] += by; d["launches"] += 1
    pk = peaks()
    traffic, traffic_src = conv_traffic()
    # bandwidth-bound kernel classes against the measured HBM copy bandwidth (algorithmic bytes / CUDA-event time)
    hbm = {k: {"achieved": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "peak": pk["gbs"], "unit": "GB/s",
               "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["gbs"], "ms_per_step": v["ms"] / prof_steps}
           for k, v in agg.items() if v["bytes"] and v["ms"] and not v["flops"]}
    conv_ms = sum(agg[k]["ms"] for k in ("conv_forward", "conv_wgrad") if k in agg)
    conv_fl = sum(agg[k]["flops"] for k in ("conv_forward", "conv_wgrad") if k in agg)
    conv_n = sum(agg[k]["launches"] for k in ("conv_forward", "conv_wgrad") if k in agg)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    breakdown = {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                     "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["flops"] and v["ms"] else None,
                     "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["bytes"] and v["ms"] else None}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = min(16, os.cpu_count() or 1)
        sample_scenes = scenes[:2]
        cpu_reference_step(sample_scenes, threads)                     # warm-up (allocator, thread pool), not timed
        ts = [cpu_reference_step(sample_scenes, threads) for _ in range(3)]
        t = float(np.median(ts))
        cpu = {"value": len(sample_scenes) / t, "unit": "scenes/s", "cores": threads, "kind": "port",
               "sample": "2 of the bench's ScanNet-shape scenes (%d voxels), fwd+loss+bwd+Adam, median of 3 runs after one "
                         "warm-up run (%s s); oracle port of ME's CPU algorithm (MinkowskiEngine 0.5.4 not installable "
                         "offline)" % (sum(len(s["vox_coords"]) for s in sample_scenes), ", ".join("%.1f" % x for x in ts))}
    total_scenes = args.scenes * world
    line = {
        "metric": METRIC, "value": total_scenes / (ms_step * 1e-3), "unit": "scenes/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs/scannet.txt training step: %d ScanNet-shape scenes/GPU (%.0f voxels/GPU at 2 cm), "
                               "hash + 16 kernel maps + fwd + box-vote losses + bwd + Adam" % (args.scenes, voxels),
                   "scenes_per_gpu": args.scenes, "voxels_per_gpu": voxels, "parallelism": "dp%d" % world,
                   "sync_bn": bool(args.sync_bn and world > 1),
                   "scene_sampling": ("size-balanced: each rank keeps the %d of %d generated scenes whose total voxel count is "
                                      "closest to the all-rank mean" % (args.scenes, args.scenes + 2)) if world > 1 else "fixed seeds",
                   "grad_sync": (args.grad_sync if world > 1 else None),
                   "timing": "median of %d regions of %d steps each, ms/step of every region: value %s, e2e %s" % (
                       REGIONS, args.steps, ["%.2f" % m for m in ms_all], ["%.2f" % m for m in ms_e2e_all]),
                   "coordinate_maps": "built one step ahead on a side stream" if args.prefetch else "built inside the step",
                   "cache": "inputs larger than L2: every full-resolution activation is >= 235 MB (L2 is 126 MB) and "
                            "every step runs on freshly translated coordinates, so all 16 kernel maps are rebuilt"},
        "e2e": {"value": total_scenes / (ms_e2e * 1e-3), "unit": "scenes/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": pk["src"] + " sustained bf16",
                     "kernel": "conv_fwd_kernel + conv_wgrad_kernel (all %d launches/step)" % (conv_n // max(prof_steps, 1)),
                     "algorithmic_flops_per_step": conv_fl / max(prof_steps, 1),
                     "kernel_ms_per_step": conv_ms / max(prof_steps, 1),
                     "timing": "CUDA event pair around every launch, one stream, %d instrumented steps after the timed "
                               "regions; the GPU is kept backlogged (maps built first, spin kernel ahead of forward and of "
                               "backward) so that a pair measures the kernel and not the host's enqueue time" % prof_steps},
        "roofline_hbm": hbm,
        "cpu_baseline": cpu,
        "kernels": breakdown,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
