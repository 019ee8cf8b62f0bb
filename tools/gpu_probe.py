"""Developer probe: run each kernel family in its own subprocess (a trap in one does not poison the CUDA
context of the others) and print error summaries. Usage on the GPU box:
    python tools/gpu_probe.py            # all probes
    python tools/gpu_probe.py conv 27 128 96 20000   # one probe in-process
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def probe_conv(kvol, c_in, c_out, n, seed=0):
    import numpy as np
    import torch
    from box2mask_b200 import ops
    from oracle import sparse_ops as so
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    dev = "cuda"
    if kvol == 1:
        nbr_np = None
        n_in = n
    else:
        n_in = n
        nbr_np = rng.integers(0, n_in, (kvol, n)).astype(np.int32)
        nbr_np[rng.random((kvol, n)) < 0.55] = -1
        if kvol >= 27:
            nbr_np[3, : n // 2] = -1          # exercise tile-level skipping
            nbr_np[5, :] = -1
    x = torch.randn(n_in, c_in)
    w = torch.randn(kvol, c_in, c_out) / np.sqrt(c_in * max(kvol * 0.45, 1))
    xb = so.bf16_round(x)
    wb = so.bf16_round(w)
    ref = so.sparse_conv(xb.double(), nbr_np, wb.double(), n_out=n)
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(dev)) if nbr_np is not None else None
    packed = ops.pack_weights(w.to(dev), 0)
    colsum = torch.zeros(2 * c_out, dtype=torch.float64, device=dev)
    y = ops.conv_forward(x.to(dev).to(torch.bfloat16), nbr, packed, kvol, n, c_out, colsum)
    torch.cuda.synchronize()
    yf = y.float().cpu().double()
    err = (yf - ref).abs()
    tol = 8e-3 * ref.abs() + 2e-2
    bad = int((err > tol).sum())
    s1 = colsum[:c_out].cpu()
    s2 = colsum[c_out:].cpu()
    e1 = float((s1 - ref.sum(0)).abs().max())
    e2 = float(((s2 - (ref * ref).sum(0)).abs() / ((ref * ref).sum(0) + 1)).max())
    print("conv fwd kvol=%d cin=%d cout=%d n=%d: max_abs_err=%.4g ref_absmax=%.3g bad=%d colsum_err=%.3g colsumsq_rel=%.3g"
          % (kvol, c_in, c_out, n, float(err.max()), float(ref.abs().max()), bad, e1, e2))
    if bad:
        idx = torch.nonzero(err > tol)[:8]
        for r, c in idx.tolist():
            print("   row %d col %d got %.5f want %.5f" % (r, c, float(yf[r, c]), float(ref[r, c])))
        rows_bad = torch.unique(torch.nonzero(err > tol)[:, 0])
        cols_bad = torch.unique(torch.nonzero(err > tol)[:, 1])
        print("   bad rows: %d (first %s)  bad cols: %d (first %s)" % (len(rows_bad), rows_bad[:10].tolist(), len(cols_bad), cols_bad[:10].tolist()))
    return bad == 0


def probe_wgrad(kvol, c_in, c_out, n, seed=0):
    import numpy as np
    import torch
    from box2mask_b200 import ops
    from oracle import sparse_ops as so
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    dev = "cuda"
    if kvol == 1:
        nbr_np = None
    else:
        nbr_np = rng.integers(0, n, (kvol, n)).astype(np.int32)
        nbr_np[rng.random((kvol, n)) < 0.55] = -1
        if kvol > 3:
            nbr_np[2, :] = -1
            nbr_np[1, 100:] = -1
    x = so.bf16_round(torch.randn(n, c_in))
    dy = so.bf16_round(torch.randn(n, c_out))
    ref = torch.zeros(kvol, c_in, c_out, dtype=torch.float64)
    if nbr_np is None:
        ref[0] = x.double().t() @ dy.double()
    else:
        for k, (i, o) in enumerate(so.map_to_pairs(nbr_np)):
            if len(i):
                ref[k] = x.double()[i].t() @ dy.double()[o]
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(dev)) if nbr_np is not None else None
    dw = ops.conv_wgrad(x.to(dev).to(torch.bfloat16), dy.to(dev).to(torch.bfloat16), nbr, kvol, n)
    torch.cuda.synchronize()
    err = (dw.cpu().double() - ref).abs()
    tol = 2e-3 * ref.abs() + 1e-3 * float(ref.abs().max()) + 1e-3
    bad = int((err > tol).sum())
    print("conv wgrad kvol=%d cin=%d cout=%d n=%d: max_abs_err=%.4g ref_absmax=%.3g bad=%d"
          % (kvol, c_in, c_out, n, float(err.max()), float(ref.abs().max()), bad))
    if bad:
        idx = torch.nonzero(err > tol)[:8]
        for k, r, c in idx.tolist():
            print("   k %d ci %d co %d got %.5f want %.5f" % (k, r, c, float(dw[k, r, c]), float(ref[k, r, c])))
        print("   bad ci:", torch.unique(torch.nonzero(err > tol)[:, 1])[:16].tolist(), " bad co:", torch.unique(torch.nonzero(err > tol)[:, 2])[:16].tolist())
    return bad == 0


PROBES = [
    ("conv", 1, 64, 64, 1000),
    ("conv", 27, 64, 64, 5000),
    ("conv", 27, 32, 32, 5000),
    ("conv", 27, 128, 96, 20000),
    ("conv", 27, 96, 96, 777),
    ("conv", 8, 256, 256, 3000),
    ("conv", 27, 512, 256, 1500),
    ("conv", 27, 256, 512, 1500),
    ("conv", 27, 384, 256, 1000),
    ("conv", 125, 16, 32, 6000),
    ("conv", 1, 32, 64, 300),
    ("conv", 27, 96, 128, 300),
    ("conv", 27, 64, 16, 300),
    ("wgrad", 1, 64, 64, 1000),
    ("wgrad", 27, 64, 64, 5000),
    ("wgrad", 27, 128, 96, 20000),
    ("wgrad", 27, 96, 96, 777),
    ("wgrad", 27, 32, 32, 5000),
    ("wgrad", 8, 256, 256, 3000),
    ("wgrad", 27, 512, 256, 1500),
    ("wgrad", 125, 16, 32, 6000),
    ("wgrad", 8, 96, 128, 100000),
]

if __name__ == "__main__":
    if len(sys.argv) > 1:
        kind = sys.argv[1]
        args = [int(a) for a in sys.argv[2:]]
        ok = probe_conv(*args) if kind == "conv" else probe_wgrad(*args)
        sys.exit(0 if ok else 1)
    failed = 0
    for p in PROBES:
        cmd = [sys.executable, os.path.abspath(__file__), p[0]] + [str(a) for a in p[1:]]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
            out = (r.stdout + r.stderr).strip().splitlines()
            print("\n".join(out[-14:]))
            if r.returncode != 0:
                failed += 1
                print("   -> FAILED rc=%d" % r.returncode)
        except subprocess.TimeoutExpired:
            failed += 1
            print("probe %s TIMED OUT" % (p,))
        sys.stdout.flush()
    print("probes failed: %d / %d" % (failed, len(PROBES)))
