"""CPU oracle for the sparse-convolution hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (box2mask_b200/) never does.

What it restates: the arithmetic that /root/reference reaches through `import MinkowskiEngine as ME`
(MinkowskiEngine==0.5.4, pinned at /root/reference/docs/installation.md:6,42). That package is NOT under
/root/reference, is not importable here and cannot be fetched (no network), and the reference ships no
tests or golden vectors for it.  ==> PARITY UNPINNED for the ME-defined semantics listed below; they are
restated from ME 0.5.4's published behaviour and cross-checked against an independent
second oracle (dense torch.nn.functional.conv3d / conv_transpose3d on a zero-filled grid,
tests/test_oracle.py):
  (i)   cross-correlation  Y[o] = sum_k X[o + delta_k] W[k]
  (ii)  offsets enumerate with the first spatial axis fastest, k = ix + K*iy + K*K*iz; odd kernels are
        centred, even kernels start at 0; offsets scale with the input tensor stride
  (iii) strided coordinates = floor(c / s) * s, de-duplicated
  (iv)  transposed conv = the strided map with in/out swapped
  (v)   kernel_volume == 1 -> plain matmul with a 2-D kernel
  (vi)  BatchNorm = torch.nn.BatchNorm1d over rows; (vii) global pooling row = batch index.
Call sites followed: models/detection_net.py:36-138,234-364, models/resnet.py:46-83,148-181.
Row order of strided maps: sorted lexicographically by (b,x,y,z) (== np.unique(axis=0)); ME leaves it
implementation-defined, so the choice is ours and the CUDA path makes the same one.

The algorithm is ME's CPU one: dictionary/sorted-key coordinate maps, then per kernel offset
index_select -> BLAS mm -> index_add_ in fp32.
"""
import numpy as np
import torch
import torch.nn.functional as F

COORD_BIAS = 32768


def pack_keys(coords):
    """int[N,4] (b,x,y,z) -> uint64 keys whose integer order is the lexicographic order."""
    c = np.asarray(coords).astype(np.int64) + COORD_BIAS
    return ((c[:, 0] << 48) | (c[:, 1] << 32) | (c[:, 2] << 16) | c[:, 3]).astype(np.uint64)


def floor_to_multiple(v, s):
    return np.floor_divide(v, s) * s


def downsample_coords(coords, new_stride):
    """(iii): unique rows of floor(c/s)*s, batch column kept; returns (out int32[M,4], parent int32[N])."""
    c = np.asarray(coords).astype(np.int64)
    q = c.copy()
    q[:, 1:] = floor_to_multiple(c[:, 1:], new_stride)
    out, inv = np.unique(q, axis=0, return_inverse=True)
    return out.astype(np.int32), inv.reshape(-1).astype(np.int32)


def kernel_offsets(ksize, tensor_stride):
    """(ii) offsets int[K^3,3], x fastest; centred for odd kernels, starting at 0 for even ones."""
    rng = np.arange(ksize) - (ksize // 2 if ksize % 2 == 1 else 0)
    offs = []
    for iz in rng:
        for iy in rng:
            for ix in rng:
                offs.append((ix, iy, iz))
    return np.asarray(offs, dtype=np.int64) * tensor_stride


def _lookup(sorted_keys, order, query_keys):
    pos = np.searchsorted(sorted_keys, query_keys)
    pos_c = np.minimum(pos, len(sorted_keys) - 1)
    hit = sorted_keys[pos_c] == query_keys
    return np.where(hit, order[pos_c], -1).astype(np.int32)


def kernel_map_submanifold(coords, tensor_stride, ksize):
    """nbr int32[K^3, N]: row of coord(o)+delta_k, or -1."""
    c = np.asarray(coords).astype(np.int64)
    keys = pack_keys(c)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    offs = kernel_offsets(ksize, tensor_stride)
    nbr = np.empty((len(offs), len(c)), dtype=np.int32)
    for k, d in enumerate(offs):
        q = c.copy()
        q[:, 1:] += d[None]
        ok = np.all((q[:, 1:] >= -COORD_BIAS) & (q[:, 1:] < COORD_BIAS), 1)
        r = _lookup(skeys, order, pack_keys(np.where(ok[:, None], q, c)))
        nbr[k] = np.where(ok, r, -1)
    return nbr


def kernel_map_submanifold_dict(coords, tensor_stride, ksize):
    """Pure-Python dictionary version of kernel_map_submanifold (small cases; validates the numpy one)."""
    table = {tuple(int(v) for v in row): i for i, row in enumerate(np.asarray(coords))}
    offs = kernel_offsets(ksize, tensor_stride)
    nbr = np.full((len(offs), len(coords)), -1, dtype=np.int32)
    for o, row in enumerate(np.asarray(coords)):
        b, x, y, z = (int(v) for v in row)
        for k, (dx, dy, dz) in enumerate(offs):
            nbr[k, o] = table.get((b, x + int(dx), y + int(dy), z + int(dz)), -1)
    return nbr


def kernel_map_stride2(fine_coords, parent, n_coarse, fine_stride):
    """k=2,s=2 maps. nbr_down int32[8,n_coarse] (child of coarse row at offset k), nbr_up int32[8,n_fine]."""
    c = np.asarray(fine_coords).astype(np.int64)
    cs = 2 * fine_stride
    o = (c[:, 1:] - floor_to_multiple(c[:, 1:], cs)) // fine_stride
    k = o[:, 0] + 2 * o[:, 1] + 4 * o[:, 2]
    n_fine = len(c)
    nbr_down = np.full((8, n_coarse), -1, dtype=np.int32)
    nbr_down[k, parent] = np.arange(n_fine, dtype=np.int32)
    nbr_up = np.full((8, n_fine), -1, dtype=np.int32)
    nbr_up[k, np.arange(n_fine)] = parent
    return nbr_down, nbr_up


def map_to_pairs(nbr):
    """ME-style kernel map: list over offsets of (in_rows, out_rows), ordered by output row."""
    pairs = []
    for k in range(nbr.shape[0]):
        out = np.nonzero(nbr[k] >= 0)[0]
        pairs.append((nbr[k][out].astype(np.int64), out.astype(np.int64)))
    return pairs


def map_triples(nbr, in_coords, out_coords):
    """Set of (k, in-coordinate, out-coordinate) triples — the order-free identity of a kernel map."""
    s = set()
    for k, (i, o) in enumerate(map_to_pairs(nbr)):
        for a, b in zip(i, o):
            s.add((k, tuple(int(v) for v in in_coords[a]), tuple(int(v) for v in out_coords[b])))
    return s


def sparse_conv(x, nbr, weight, n_out=None):
    """ME's CPU algorithm: for each offset, gather input rows, mm with W[k], scatter-add (fp32, autograd-able).
    x [N_in, C_in], nbr [K, N_out], weight [K, C_in, C_out] (or [C_in, C_out] when K == 1 and nbr is None)."""
    if nbr is None:
        w = weight if weight.dim() == 2 else weight[0]
        return x @ w
    n_out = nbr.shape[1] if n_out is None else n_out
    y = x.new_zeros((n_out, weight.shape[-1]))
    for k, (i, o) in enumerate(map_to_pairs(np.asarray(nbr))):
        if len(i) == 0:
            continue
        ii, oo = torch.from_numpy(i), torch.from_numpy(o)
        y = y.index_add(0, oo, x.index_select(0, ii) @ weight[k])
    return y


def segment_mean(f, ids, s):
    out = f.new_zeros((s, f.shape[1])).index_add(0, ids, f)
    cnt = torch.bincount(ids, minlength=s).to(f.dtype).clamp(min=1)
    return out / cnt[:, None]


def segment_max(f, ids, s):
    out = f.new_full((s, f.shape[1]), float("-inf"))
    return out.scatter_reduce(0, ids[:, None].expand_as(f), f, reduce="amax", include_self=True)


def batch_norm(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5):
    return F.batch_norm(x, running_mean, running_var, weight, bias, training, momentum, eps)


def bf16_round(t):
    """Round-to-nearest-even to bfloat16 and back (emulates the CUDA path's storage precision)."""
    return t.to(torch.bfloat16).to(torch.float32)


# ------------------------------------------------------------------------------------------------
# second, independent oracle: densify and use torch's dense convolutions
# ------------------------------------------------------------------------------------------------
def dense_conv_reference(coords, x, weight, ksize, stride, transposed, out_coords):
    """Dense cross-check for batch-0-only coordinates with tensor stride 1 input (or stride-2 input for
    the transposed case). Returns rows at out_coords, in fp64."""
    c = np.asarray(coords).astype(np.int64)
    oc = np.asarray(out_coords).astype(np.int64)
    assert np.all(c[:, 0] == 0) and np.all(oc[:, 0] == 0)
    cin, cout = weight.shape[-2], weight.shape[-1]
    kvol = ksize ** 3
    w = weight.reshape(kvol, cin, cout).double()
    # W_dense[co, ci, dz, dy, dx] for conv3d on a grid indexed [z, y, x]
    wd = w.reshape(ksize, ksize, ksize, cin, cout).permute(4, 3, 0, 1, 2).contiguous()
    if not transposed:
        pad = ksize // 2 if ksize % 2 == 1 else 0
        size = int(max(c[:, 1:].max(), oc[:, 1:].max())) + ksize + 2
        size += size % 2
        grid = torch.zeros((1, cin, size, size, size), dtype=torch.float64)
        grid[0, :, c[:, 3], c[:, 2], c[:, 1]] = x.double().t()
        out = F.conv3d(grid, wd, stride=stride, padding=pad)
        return out[0][:, oc[:, 3] // stride, oc[:, 2] // stride, oc[:, 1] // stride].t()
    # transposed k=2 s=2: input lives on even coordinates (tensor stride 2)
    size = int(max(c[:, 1:].max() // 2, oc[:, 1:].max() // 2)) + 2
    grid = torch.zeros((1, cin, size, size, size), dtype=torch.float64)
    grid[0, :, c[:, 3] // 2, c[:, 2] // 2, c[:, 1] // 2] = x.double().t()
    wt = w.reshape(ksize, ksize, ksize, cin, cout).permute(3, 4, 0, 1, 2).contiguous()  # [ci, co, dz, dy, dx]
    out = F.conv_transpose3d(grid, wt, stride=2)
    return out[0][:, oc[:, 3], oc[:, 2], oc[:, 1]].t()
