import ctypes, os, sys
import numpy as np, torch
ROOT="/root/repo"; sys.path.insert(0, ROOT)
from box2mask_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "libb2m_dbg.so")
from box2mask_b200 import ops
lib=_lib.load(); lib.b2m_debug_wait_buffer.restype = ctypes.POINTER(ctypes.c_uint)
torch.zeros(1, device="cuda"); buf = lib.b2m_debug_wait_buffer()
kvol,cin,cout,n = [int(v) for v in sys.argv[1:5]]
rng=np.random.default_rng(kvol+cin); torch.manual_seed(kvol+cin)
nbr_np = rng.integers(0, n, (kvol, n)).astype(np.int32); nbr_np[rng.random((kvol, n)) < 0.55] = -1
km = ops.sort_kernel_map(torch.from_numpy(nbr_np).cuda())
x = torch.randn(n, cin, device="cuda").to(torch.bfloat16); w = torch.randn(kvol, cin, cout, device="cuda")*0.05
try:
    y = ops.conv_forward(x, km, ops.pack_weights(w, 0), kvol, n, cout); torch.cuda.synchronize(); print("ok")
except Exception as e: print("fail", str(e)[:80])
cnt=buf[0]; print("timeouts", cnt)
for i in range(min(cnt,40)):
    r=[buf[8+i*8+j] for j in range(6)]
    print("block (%d,%d) warp %2d bar+%d parity %d tag %d" % (r[0], r[5], r[1]//32, r[2]&0xFFFF, r[3], r[4]))
