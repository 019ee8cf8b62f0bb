"""Ablation of the forward convolution kernel with the debug build (build/libb2m_dbg.so): B2M_ABLATE bit 0 = no A
gathers, 1 = no MMAs, 2 = no epilogue stores/statistics, 3 = no B copies. Shows which role bounds the kernel."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
from box2mask_b200 import _lib
_lib.LIB_PATH = os.path.join(%r, "build", "libb2m_dbg.so")
sys.argv = ["conv_bench.py"] + sys.argv[1:]
__file__ = os.path.join(%r, "tools", "conv_bench.py")
exec(open(__file__).read())
''' % (ROOT, ROOT, ROOT)
for ab in (0, 1, 2, 4, 8, 9, 3, 15):
    env = dict(os.environ, B2M_ABLATE=str(ab))
    r = subprocess.run([sys.executable, "-c", code] + sys.argv[1:], env=env, capture_output=True, text=True, timeout=120)
    line = [l for l in r.stdout.splitlines() if l.startswith("fwd")]
    print("ablate=%2d  %s" % (ab, line[-1] if line else (r.stderr.strip().splitlines() or ["?"])[-1][:100]))
