"""Run GPU tests against the DEBUG build and, if a kernel trapped on a timed-out barrier wait, print which waits timed
out (block, warp, barrier offset, parity, tag).   python tools/dbg_pytest.py tests/test_gpu_kernels.py -k real_maps -x"""
import ctypes
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B2M_LIB", os.path.join(ROOT, "build", "libb2m_dbg.so"))
import pytest  # noqa: E402
import torch  # noqa: E402
from box2mask_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.b2m_debug_wait_buffer.restype = ctypes.POINTER(ctypes.c_uint)
torch.zeros(1, device="cuda")
buf = lib.b2m_debug_wait_buffer()
rc = pytest.main(["-q", "-m", "gpu"] + sys.argv[1:])
cnt = buf[0]
print("pytest rc", rc, "timed-out waits:", cnt)
rows = [tuple(buf[8 + i * 8 + j] for j in range(6)) for i in range(min(cnt, 255))]
print("by (warp, tag):", sorted(Counter((r[1] // 32, r[4]) for r in rows).items()))
for r in rows[:40]:
    print("block (%3d,%d) warp %2d lane %2d bar+%4d parity %d tag %d" % (r[0], r[5], r[1] // 32, r[1] % 32, r[2] & 0xFFFF, r[3], r[4]))
