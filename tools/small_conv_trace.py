"""Where does a small (deep-level) convolution launch spend its time? Runs conv forward / wgrad of a 256-wide k=3
layer on the real kernel maps of the deep levels of a ScanNet-shape batch with the DEBUG build (build/libb2m_dbg.so,
python tools/build_variant.py dbg -DB2M_DEBUG_BUILD), which stamps the SM clock of block (0,0,0) at fixed points of
the kernels (B2M_TRACE in csrc/conv.cu), and prints the stamps in microseconds relative to kernel entry, next to the
CUDA-event time of the launch.   python tools/small_conv_trace.py [--lib build/libb2m_dbg.so]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2mask_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "build", "libb2m_dbg.so"))
ap.add_argument("--scenes", type=int, default=8)
ap.add_argument("--mhz", type=float, default=1965.0)
ap.add_argument("--levels", default="3,4,5,6,7")
ap.add_argument("--channels", type=int, default=256)
ap.add_argument("--which", default="fwd,wgrad")
args = ap.parse_args()
_lib.LIB_PATH = args.lib
from box2mask_b200 import ops  # noqa: E402
from box2mask_b200.me.sparse_tensor import CoordinateManager  # noqa: E402
from box2mask_b200.synthetic import batched_coordinates, make_scene  # noqa: E402

lib = _lib.load()
buf = None
if hasattr(lib, "b2m_debug_wait_buffer"):
    lib.b2m_debug_wait_buffer.restype = ctypes.POINTER(ctypes.c_uint)
    torch.zeros(1, device="cuda")
    buf = lib.b2m_debug_wait_buffer()

coords = batched_coordinates([make_scene(10000 + i, scale=0.84)["vox_coords"] for i in range(args.scenes)]).cuda()
cm = CoordinateManager(coords)
cm.prepare(7, [(2 ** l, 3) for l in range(8)])
C = args.channels


def stamps():
    if buf is None:
        return {}
    v = {i: buf[4096 + i] for i in range(160) if buf[4096 + i]}
    if 0 not in v:
        return {}
    t0 = v[0]
    return {i: ((t - t0) & 0xFFFFFFFF) / args.mhz for i, t in sorted(v.items())}


def clear():
    if buf is not None:
        for i in range(160):
            buf[4096 + i] = 0


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    # back-to-back launches: the per-launch time when the GPU never waits for the host
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) * 1e3 / 20
    # the same 20 launches replayed from a CUDA graph: pure GPU time per launch (no host dispatch in between)
    gms = float("nan")
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    fn()
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) * 1e3 / 20
    except Exception as e:  # noqa: BLE001
        print("   (graph capture failed: %s)" % str(e)[:80])
    clear(); fn(); torch.cuda.synchronize()
    return best, b2b, gms, stamps()


for lvl in [int(v) for v in args.levels.split(",")]:
    km = cm.submanifold_map(2 ** lvl, 3)
    n = cm.coords(2 ** lvl).shape[0]
    x = torch.randn(n, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(27, C, C, device="cuda") * 0.05
    packed = ops.pack_weights(w, 0)
    colsum = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    for name, fn in (("fwd", lambda: ops.conv_forward(x, km, packed, 27, n, C, colsum)),
                     ("wgrad", lambda: ops.conv_wgrad(x, dy, km, 27, n))):
        if name not in args.which.split(","):
            continue
        single, b2b, gms, st = timed(fn)
        print("L%d n=%d %s %d->%d: single launch %.1f us, back-to-back %.1f us/launch, in a CUDA graph %.1f us/launch" % (
            lvl, n, name, C, C, single, b2b, gms))
        if st:
            print("   stamps(us): " + "  ".join("%d:%.1f" % (k, v) for k, v in st.items()))
