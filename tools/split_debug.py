"""Offset-split vs unsplit convolution forward against the fp64 oracle (debugging aid): error statistics of both paths."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from box2mask_b200 import _lib, ops  # noqa: E402
from oracle import sparse_ops as so  # noqa: E402
from test_gpu_kernels import _conv_case  # noqa: E402

DEV = "cuda"
for kvol, c_in, c_out, n in [(27, 96, 96, 515), (27, 256, 256, 700), (8, 128, 96, 3000)]:
    nbr_np, x, w, n = _conv_case(kvol, c_in, c_out, n, seed=3)
    ref = so.sparse_conv(x.double(), nbr_np, w.double(), n_out=n)
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV))
    xd, wp = x.to(DEV).to(torch.bfloat16), ops.pack_weights(w.to(DEV), 0)
    out = {}
    for name, opts in (("split", {}), ("unsplit", {_lib.OPT_SPLIT_OFFSETS: 1}), ("unsplit_general", {_lib.OPT_SPLIT_OFFSETS: 1, _lib.OPT_ISSUER: 1}),
                       ("unsplit_tma", {_lib.OPT_SPLIT_OFFSETS: 1, _lib.OPT_GATHER_MODE: 1}), ("split_tma", {_lib.OPT_GATHER_MODE: 1})):
        for k, v in opts.items():
            _lib.set_option(k, v)
        y = ops.conv_forward(xd, nbr, wp, kvol, n, c_out).float().cpu().double()
        for k in opts:
            _lib.set_option(k, 0)
        out[name] = y
        err = (y - ref).abs()
        ulp = 2.0 ** (torch.floor(torch.log2(ref.abs().clamp(min=1e-30))) - 7)
        print("k%d %d->%d n=%d %-16s max|err| %.5f  max err/ulp %.3f  frac(err > 0.51 ulp) %.5f" % (
            kvol, c_in, c_out, n, name, float(err.max()), float((err / ulp).max()), float((err > 0.51 * ulp).double().mean())))
    d = (out["split"] - out["unsplit"]).abs()
    i = int(d.argmax())
    r, c = i // c_out, i % c_out
    print("   split vs unsplit: differ in %.4f of the elements, max %.5f at (%d,%d): split %.6f unsplit %.6f ref %.6f" % (
        float((d > 0).double().mean()), float(d.max()), r, c, float(out["split"][r, c]), float(out["unsplit"][r, c]), float(ref[r, c])))
