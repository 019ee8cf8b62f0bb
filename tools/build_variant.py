"""Build a variant of the C-ABI library into build/libb2m_<name>.so (git-ignored, travels with gpurun):
    python tools/build_variant.py dbg -DB2M_DEBUG_BUILD
Load it with  box2mask_b200._lib.LIB_PATH = ".../build/libb2m_<name>.so"  before the first op (tools/ablate_bench.py,
tools/debug_wait.py, tools/small_conv_trace.py)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2mask_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
objdir = os.path.join(ROOT, "build", name)
os.makedirs(objdir, exist_ok=True)


def one(src):
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr)
    return obj


with ThreadPoolExecutor(max_workers=5) as ex:
    objs = list(ex.map(one, B.SOURCES))
out = os.path.join(ROOT, "build", "libb2m_%s.so" % name)
subprocess.check_call([B._nvcc(), "-shared", "-o", out] + objs + ["-lcudart"])
print(out)
