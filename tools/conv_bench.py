"""Micro-benchmark of the two convolution kernels on a REAL kernel map (k3 at full resolution of a synthetic
ScanNet-shape batch). For ncu: python tools/conv_bench.py --scenes 8 --cin 128 --cout 96 --iters 3"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("B2M_BENCH_LIB"):      # a variant built by tools/build_variant.py
    from box2mask_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(os.environ["B2M_BENCH_LIB"])
from box2mask_b200 import ops  # noqa: E402
from box2mask_b200.synthetic import batched_coordinates, make_scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=8)
ap.add_argument("--scale", type=float, default=0.84)
ap.add_argument("--cin", type=int, default=128)
ap.add_argument("--cout", type=int, default=96)
ap.add_argument("--ksize", type=int, default=3)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--block-rows", type=int, default=None)
ap.add_argument("--which", default="fwd,wgrad")
ap.add_argument("--gather", default="both", help="comma list of cpasync | tma | cpasync2 (b2m_set_option B2M_OPT_GATHER_MODE); both = cpasync,tma")
ap.add_argument("--issuer", default="lean", help="comma list of lean | general (B2M_OPT_ISSUER)")
ap.add_argument("--wgrows", default="0", help="comma list of 0 | 64 (B2M_OPT_WGRAD_ROWS)")
ap.add_argument("--bslots", default="0", help="comma list of dY ring depths (B2M_OPT_WGRAD_BSLOTS)")
ap.add_argument("--wggroup", default="0", help="comma list of 0 | 2 (B2M_OPT_WGRAD_GROUP)")
args = ap.parse_args()

dev = "cuda"
cache = "/tmp/conv_bench_coords_%d_%.2f.npy" % (args.scenes, args.scale)
if os.path.exists(cache):
    coords_np = np.load(cache)
else:
    coords_np = batched_coordinates([make_scene(10000 + i, scale=args.scale)["vox_coords"] for i in range(args.scenes)]).numpy()
    np.save(cache, coords_np)
coords = torch.from_numpy(coords_np).to(dev)
n = coords.shape[0]
t0 = time.time()
table = ops.hash_build(coords)
nbr = ops.kernel_map_submanifold(coords, 1, args.ksize, table)
km = ops.sort_kernel_map(nbr, n, block_rows=args.block_rows)
torch.cuda.synchronize()
kvol = args.ksize ** 3
pairs = int((nbr >= 0).sum())
gm = km.gmask.cpu().numpy().view(np.uint32)
bits = np.unpackbits(gm.view(np.uint8).reshape(gm.shape[0], -1), axis=1, bitorder="little")[:, :kvol]
print("rows %d pairs/row %.2f  non-empty (64-row group, offset) blocks: %.3f  density in them: %.3f" % (
    n, pairs / n, bits.mean(), pairs / (bits.sum() * 64.0)))
x = torch.randn(n, args.cin, device=dev).to(torch.bfloat16)
dy = torch.randn(n, args.cout, device=dev).to(torch.bfloat16)
w = torch.randn(kvol, args.cin, args.cout, device=dev) * 0.05
packed = ops.pack_weights(w, 0)
flops = 2.0 * pairs * args.cin * args.cout


def bench(fn, name):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    print("%-30s k%d %d->%d rows %d: %.3f ms  %.1f TFLOP/s (algorithmic)" % (name, kvol, args.cin, args.cout, n, ms, flops / ms / 1e9))


from box2mask_b200 import _lib as L  # noqa: E402
_set = L.set_option


def _tolerant(opt, val):      # older library variants do not know the newer options
    try:
        _set(opt, val)
    except Exception:
        pass


L.set_option = _tolerant
GM = {"cpasync": 0, "tma": 1, "cpasync2": 2, "cpasync_all": 3}
for mode in (["cpasync", "tma"] if args.gather == "both" else args.gather.split(",")):
    L.set_option(L.OPT_GATHER_MODE, GM[mode])
    if "fwd" in args.which:
        for iss in args.issuer.split(","):
            L.set_option(L.OPT_ISSUER, 1 if iss == "general" else 0)
            colsum = torch.zeros(2 * args.cout, dtype=torch.float64, device=dev)
            bench(lambda: ops.conv_forward(x, km, packed, kvol, n, args.cout, colsum), "fwd[%s,%s]" % (mode, iss))
        L.set_option(L.OPT_ISSUER, 0)
    if "dgrad" in args.which:
        # the dgrad of a unit (square layer: cin == cout here) with the pending-gradient add, without / with the
        # BatchNorm-backward reduction of the producer layer in its epilogue, and that reduction as a pass of its own
        pend = torch.randn(n, args.cout, device=dev).to(torch.bfloat16)
        px = torch.randn(n, args.cout, device=dev).to(torch.bfloat16)
        mean, invstd = torch.randn(args.cout, device=dev), torch.rand(args.cout, device=dev) + 0.5
        mask = torch.randint(0, 256, (n, args.cout // 8), dtype=torch.uint8, device=dev)
        red = torch.zeros(2 * args.cout, dtype=torch.float64, device=dev)
        bench(lambda: ops.conv_forward(x, km, packed, kvol, n, args.cout), "[%s] no statistics, no residual" % mode)
        bench(lambda: ops.conv_forward(x, km, packed, kvol, n, args.cout, residual=pend), "dgrad[%s] plain" % mode)
        bench(lambda: ops.conv_forward(x, km, packed, kvol, n, args.cout, residual=pend, bn_reduce=(px, mask, mean, invstd, red)),
              "dgrad[%s] + BN reduction" % mode)
        g = ops.conv_forward(x, km, packed, kvol, n, args.cout, residual=pend)
        bench(lambda: L.check(L.load().b2m_bn_backward_reduce(L.ptr(px), None, L.ptr(g), n, args.cout, L.ptr(mean), L.ptr(invstd),
                                                               1, L.ptr(red), L.ptr(mask), L.stream_ptr()), "reduce"),
              "bn_backward_reduce alone")
    if "wgrad" in args.which:
        for r in args.wgrows.split(","):
            L.set_option(L.OPT_WGRAD_ROWS, int(r))
            for gsz in args.wggroup.split(","):
                L.set_option(L.OPT_WGRAD_GROUP, int(gsz))
                for bs in args.bslots.split(","):
                    L.set_option(L.OPT_WGRAD_BSLOTS, int(bs))
                    bench(lambda: ops.conv_wgrad(x, dy, km, kvol, n), "wgrad[%s,rows%s,grp%s,bs%s]" % (mode, r, gsz, bs))
                L.set_option(L.OPT_WGRAD_BSLOTS, 0)
            L.set_option(L.OPT_WGRAD_GROUP, 0)
        L.set_option(L.OPT_WGRAD_ROWS, 0)
