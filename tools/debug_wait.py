"""Debug aid: run one conv forward with the debug build (build/libb2m_dbg.so, -DB2M_DEBUG_BUILD) and print which
barrier waits timed out (block, warp, barrier offset, parity, tag) from the mapped host buffer."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2mask_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "build", "libb2m_dbg.so")
from box2mask_b200 import ops  # noqa: E402
from box2mask_b200.me.utils import batched_coordinates  # noqa: E402
from box2mask_b200.synthetic import make_scene  # noqa: E402

lib = _lib.load()
lib.b2m_debug_wait_buffer.restype = ctypes.POINTER(ctypes.c_uint)
torch.zeros(1, device="cuda")
buf = lib.b2m_debug_wait_buffer()
cin, cout, scenes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
coords = batched_coordinates([make_scene(10000 + i, scale=0.84)["vox_coords"] for i in range(scenes)]).cuda()
n = coords.shape[0]
nbr = ops.kernel_map_submanifold(coords, 1, 3, ops.hash_build(coords))
km = ops.sort_kernel_map(nbr, n)
x = torch.randn(n, cin, device="cuda").to(torch.bfloat16)
w = torch.randn(27, cin, cout, device="cuda") * 0.05
packed = ops.pack_weights(w, 0)
try:
    for _ in range(20):
        y = ops.conv_forward(x, km, packed, 27, n, cout)
    torch.cuda.synchronize()
    print("no failure; rows", n)
except Exception as e:  # noqa: BLE001
    print("failure:", str(e)[:120])
cnt = buf[0]
print("timed-out waits:", cnt)
rows = [tuple(buf[8 + i * 8 + j] for j in range(6)) for i in range(min(cnt, 255))]
from collections import Counter  # noqa: E402
print("by (warp, tag):", sorted(Counter((r[1] // 32, r[4]) for r in rows).items()))
for r in rows[:24]:
    print("block %3d warp %2d bar+%4d parity %d tag %d" % (r[0], r[1] // 32, r[2] & 0xFFFF, r[3], r[4]))
