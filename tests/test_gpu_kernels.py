"""GPU parity tests (run with -m gpu on the B200 box): every C-ABI kernel against the CPU oracle on the
same seeded inputs. Integer/index work must be bit-exact; floating point within the stated tolerance."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from box2mask_b200 import ops  # noqa: E402
from box2mask_b200.synthetic import make_boxes, make_scene  # noqa: E402
from oracle import nms as onms  # noqa: E402
from oracle import sparse_ops as so  # noqa: E402

DEV = "cuda"


def _coords(xyz, b=0):
    xyz = np.asarray(xyz, dtype=np.int32)
    return np.concatenate([np.full((len(xyz), 1), b, np.int32), xyz], 1)


@pytest.fixture(scope="module")
def scene_coords():
    s0 = make_scene(3, scale=0.45)
    s1 = make_scene(4, scale=0.35)
    return np.concatenate([_coords(s0["vox_coords"], 0), _coords(s1["vox_coords"], 1)], 0)


# ------------------------------------------------------------------------------------------------
# coordinates, hash, kernel maps: bit-exact
# ------------------------------------------------------------------------------------------------
def test_hash_build_query(scene_coords):
    c = torch.from_numpy(scene_coords).to(DEV)
    table = ops.hash_build(c)
    assert table.status.cpu().tolist() == [0, 0]
    rows = ops.hash_query(table, c)
    assert torch.equal(rows.cpu(), torch.arange(len(c), dtype=torch.int32))
    q = c.clone()
    q[:, 1] += 1000
    assert int((ops.hash_query(table, q) >= 0).sum()) == 0
    # duplicates keep the first occurrence; out-of-range coordinates are counted, not inserted
    d = torch.cat([c[:100], c[:50], torch.tensor([[0, 40000, 0, 0]], dtype=torch.int32, device=DEV)], 0)
    t2 = ops.hash_build(d)
    assert t2.status.cpu().tolist() == [50, 1]
    assert torch.equal(ops.hash_query(t2, c[:100]).cpu(), torch.arange(100, dtype=torch.int32))


def test_empty_inputs():
    c = torch.zeros((0, 4), dtype=torch.int32, device=DEV)
    t = ops.hash_build(c)
    out, parent = ops.downsample_coords(c, 2)
    assert out.shape == (0, 4) and parent.shape == (0,)
    nbr = ops.kernel_map_submanifold(c, 1, 3, t)
    assert nbr.shape == (27, 0)


@pytest.mark.parametrize("ksize,ts", [(3, 1), (5, 1)])
def test_kernel_map_submanifold_bit_exact(scene_coords, ksize, ts):
    c = torch.from_numpy(scene_coords).to(DEV)
    nbr = ops.kernel_map_submanifold(c, ts, ksize, ops.hash_build(c))
    ref = so.kernel_map_submanifold(scene_coords, ts, ksize)
    n = len(scene_coords)
    assert nbr.shape == (ksize ** 3, ops.map_pitch(n))          # padded pitch, -1 in the padding
    assert np.array_equal(nbr.cpu().numpy()[:, :n], ref) and bool((nbr[:, n:] == -1).all())
    counts = ops.kernel_map_count(nbr, n).cpu().numpy()
    assert np.array_equal(counts, (ref >= 0).sum(1))


def test_strided_levels_bit_exact(scene_coords):
    """All 7 stride-2 levels: coordinates (sorted unique), parent rows, k2s2 maps and the k3 map at each level."""
    cur_np, cur = scene_coords, torch.from_numpy(scene_coords).to(DEV)
    ts = 1
    for level in range(7):
        out, parent = ops.downsample_coords(cur, 2 * ts)
        ref_out, ref_parent = so.downsample_coords(cur_np, 2 * ts)
        assert np.array_equal(out.cpu().numpy(), ref_out), level
        assert np.array_equal(parent.cpu().numpy(), ref_parent), level
        nd, nu = ops.kernel_map_stride2(cur, parent, len(ref_out), ts)
        rd, ru = so.kernel_map_stride2(cur_np, ref_parent, len(ref_out), ts)
        assert np.array_equal(nd.cpu().numpy()[:, :len(ref_out)], rd), level
        assert np.array_equal(nu.cpu().numpy()[:, :len(cur_np)], ru), level
        assert bool((nd[:, len(ref_out):] == -1).all()) and bool((nu[:, len(cur_np):] == -1).all()), level
        ts *= 2
        cur_np, cur = ref_out, out
        nbr = ops.kernel_map_submanifold(cur, ts, 3, ops.hash_build(cur))
        assert np.array_equal(nbr.cpu().numpy()[:, :len(cur_np)], so.kernel_map_submanifold(cur_np, ts, 3)), level


def test_hierarchical_kernel_maps_bit_exact(scene_coords):
    """CoordinateManager.prepare builds only the coarsest level's 3^3 map by hashing; every finer level's map (and the
    5^3 map of the input level) is derived from the coarser level's tables (b2m_kernel_map_from_coarse). All of them
    must equal the oracle's tables entry by entry."""
    from box2mask_b200.me.sparse_tensor import CoordinateManager
    for coords_np in (scene_coords, None):
        if coords_np is None:      # negative coordinates, unsorted input rows
            rng = np.random.default_rng(9)
            xyz = np.unique(rng.integers(-40, 40, (6000, 3)), axis=0)
            coords_np = _coords(xyz)[rng.permutation(len(xyz))]
        cm = CoordinateManager(torch.from_numpy(coords_np).to(DEV))
        cm.prepare(7, [(1, 5)] + [(2 ** l, 3) for l in range(8)])
        assert not cm.raw_k3 and len(cm.tables) == 1 and 128 in cm.tables     # one hash table: the coarsest level
        cur = coords_np
        for l in range(8):
            ts = 2 ** l
            assert np.array_equal(cm.levels[ts].cpu().numpy(), cur), l
            for k in ((3, 5) if l == 0 else (3,)):
                km = cm.sub_maps[(ts, k)]
                ref = so.kernel_map_submanifold(cur, ts, k)
                n = len(cur)
                got = km.nbr.cpu().numpy()[:, :n]
                if km.order is not None:
                    order = km.order.cpu().numpy()[:n]
                    unsorted = np.empty_like(got)
                    unsorted[:, order] = got
                    got = unsorted
                assert np.array_equal(got, ref), (l, k)
                assert bool((km.nbr[:, n:] == -1).all())
                if k == 5:       # group masks of the unsorted 125-offset table, produced by the same kernel
                    g = (n + 63) // 64
                    bits = np.unpackbits(km.gmask.cpu().numpy().view(np.uint8).reshape(g, 16), axis=1, bitorder="little")[:, :125]
                    ref_bits = np.pad(ref >= 0, ((0, 0), (0, g * 64 - n))).reshape(125, g, 64).any(2).T
                    assert np.array_equal(bits.astype(bool), ref_bits)
            if l < 7:
                cur, _ = so.downsample_coords(cur, 2 * ts)


@pytest.mark.parametrize("block_rows", [4096, 0])
def test_sorted_kernel_map(scene_coords, block_rows):
    """b2m_kernel_map_sort: a permutation of the rows, stable inside (block, mask) classes; the pair set is unchanged."""
    nbr_np = so.kernel_map_submanifold(scene_coords, 1, 3)
    n = nbr_np.shape[1]
    km = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV), block_rows=block_rows)
    assert km.order.shape == (ops.map_pitch(n),) and km.nbr.shape == (27, ops.map_pitch(n))
    assert bool((km.order[n:] == -1).all()) and bool((km.nbr[:, n:] == -1).all())
    order = km.order.cpu().numpy()[:n]
    assert np.array_equal(np.sort(order), np.arange(n))
    assert np.array_equal(km.nbr.cpu().numpy()[:, :n], nbr_np[:, order])
    mask = np.zeros(n, np.uint64)
    for k in range(27):
        mask |= (nbr_np[k] >= 0).astype(np.uint64) << np.uint64(k)
    if block_rows:
        ref_order = np.lexsort((np.arange(n), mask, np.arange(n) // block_rows))
        assert np.array_equal(order, ref_order)
    else:
        assert np.array_equal(order, np.arange(n))
    g = (n + 63) // 64
    pad = g * 64 - n
    gm = np.bitwise_or.reduce(np.pad(mask[order], (0, pad)).reshape(g, 64), axis=1).astype(np.uint32)
    assert np.array_equal(km.gmask.cpu().numpy().view(np.uint32)[:, 0], gm)
    # 125-offset map: never sorted, 4 mask words
    nbr5 = so.kernel_map_submanifold(scene_coords[:3000], 1, 5)
    km5 = ops.sort_kernel_map(torch.from_numpy(nbr5).to(DEV))
    assert np.array_equal(km5.order.cpu().numpy()[:3000], np.arange(3000)) and km5.gmask.shape == (47, 4)
    bits = np.unpackbits(km5.gmask.cpu().numpy().view(np.uint8).reshape(47, 16), axis=1, bitorder="little")[:, :125]
    ref_bits = np.pad(nbr5 >= 0, ((0, 0), (0, 47 * 64 - 3000))).reshape(125, 47, 64).any(2).T
    assert np.array_equal(bits.astype(bool), ref_bits)


def test_kernel_map_negative_and_unsorted():
    rng = np.random.default_rng(5)
    xyz = np.unique(rng.integers(-20, 20, (3000, 3)), axis=0)
    c_np = _coords(xyz)[rng.permutation(len(xyz))]
    c = torch.from_numpy(c_np).to(DEV)
    nbr = ops.kernel_map_submanifold(c, 1, 3, ops.hash_build(c))
    assert np.array_equal(nbr.cpu().numpy()[:, :len(c_np)], so.kernel_map_submanifold(c_np, 1, 3))
    out, parent = ops.downsample_coords(c, 2)
    ro, rp = so.downsample_coords(c_np, 2)
    assert np.array_equal(out.cpu().numpy(), ro) and np.array_equal(parent.cpu().numpy(), rp)


# ------------------------------------------------------------------------------------------------
# convolution: bf16 operands, fp32 accumulation.  Tolerance: |err| <= 8e-3*|ref| + 2e-2 against an fp64
# evaluation on the same bf16-rounded operands (bf16 output rounding is 2^-9 relative).
# ------------------------------------------------------------------------------------------------
def _conv_case(kvol, c_in, c_out, n, seed, real_map=None):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    if real_map is not None:
        nbr_np = real_map
        n = nbr_np.shape[1]
    elif kvol == 1:
        nbr_np = None
    else:
        nbr_np = rng.integers(0, n, (kvol, n)).astype(np.int32)
        nbr_np[rng.random((kvol, n)) < 0.55] = -1
    x = so.bf16_round(torch.randn(n, c_in))
    w = so.bf16_round(torch.randn(kvol, c_in, c_out) / np.sqrt(c_in * max(kvol * 0.45, 1)))
    return nbr_np, x, w, n


@pytest.mark.parametrize("kvol,c_in,c_out,n", [
    (27, 32, 32, 4000), (27, 64, 64, 3000), (27, 128, 96, 6000), (27, 96, 96, 515), (8, 256, 256, 2000),
    (27, 512, 256, 700), (27, 256, 512, 700), (27, 384, 256, 500), (125, 16, 32, 3000), (1, 32, 64, 300),
    (1, 256, 256, 129), (27, 192, 128, 1), (8, 96, 96, 127), (8, 48, 32, 500), (27, 80, 64, 300), (27, 160, 16, 260),
    (1, 96, 128, 40000), (27, 32, 96, 33000), (125, 8, 32, 3000), (27, 8, 16, 200), (1, 8, 32, 90000),
])
def test_conv_forward(kvol, c_in, c_out, n):
    nbr_np, x, w, n = _conv_case(kvol, c_in, c_out, n, seed=kvol + c_in)
    ref = so.sparse_conv(x.double(), nbr_np, w.double(), n_out=n)
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV)) if nbr_np is not None else None
    colsum = torch.zeros(2 * c_out, dtype=torch.float64, device=DEV)
    y = ops.conv_forward(x.to(DEV).to(torch.bfloat16), nbr, ops.pack_weights(w.to(DEV), 0), kvol, n, c_out, colsum)
    err = (y.float().cpu().double() - ref).abs()
    assert bool((err <= 8e-3 * ref.abs() + 2e-2).all()), float(err.max())
    assert torch.allclose(colsum[:c_out].cpu(), ref.sum(0), rtol=1e-4, atol=1e-2)
    assert torch.allclose(colsum[c_out:].cpu(), (ref * ref).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("kvol,c_in,c_out,n", [(27, 256, 256, 700), (27, 96, 96, 515), (8, 128, 96, 3000), (125, 8, 32, 3000)])
def test_conv_forward_offset_split_equals_unsplit(kvol, c_in, c_out, n):
    """Few row tiles: the offsets are split over CTAs and summed by the finalize kernel in a fixed order. Same inputs with
    the split switched off (b2m_set_option) must agree to fp32 summation-order noise, and the split run is deterministic."""
    from box2mask_b200 import _lib
    nbr_np, x, w, n = _conv_case(kvol, c_in, c_out, n, seed=3)
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV))
    xd, wp = x.to(DEV).to(torch.bfloat16), ops.pack_weights(w.to(DEV), 0)
    assert _lib.load().b2m_conv_forward_workspace_bytes(n, c_in, kvol, c_out) > 0
    cs1 = torch.zeros(2 * c_out, dtype=torch.float64, device=DEV)
    y1 = ops.conv_forward(xd, nbr, wp, kvol, n, c_out, cs1)
    y1b = ops.conv_forward(xd, nbr, wp, kvol, n, c_out)
    assert torch.equal(y1, y1b)
    _lib.set_option(_lib.OPT_SPLIT_OFFSETS, 1)
    try:
        assert _lib.load().b2m_conv_forward_workspace_bytes(n, c_in, kvol, c_out) == 0
        cs0 = torch.zeros(2 * c_out, dtype=torch.float64, device=DEV)
        y0 = ops.conv_forward(xd, nbr, wp, kvol, n, c_out, cs0)
    finally:
        _lib.set_option(_lib.OPT_SPLIT_OFFSETS, 0)
    d = (y1.float() - y0.float()).abs()
    big = torch.maximum(y0.float().abs(), y1.float().abs())
    # at most one bf16 ulp apart (a rounding tie broken the other way by fp32 summation order); outputs that cancel to
    # ~0 differ by that summation noise itself (measured <= 1e-5 absolute, many ulps of a tiny value): 1e-4 absolute slack
    assert bool((d <= 2 ** -7 * big + 1e-4).all()), float(d.max())
    assert float((d > 0).float().mean()) < 0.02
    assert torch.allclose(cs1, cs0, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("kvol,c_in,c_out,n,relu", [(27, 96, 96, 515, True), (27, 96, 128, 40000, True), (1, 96, 96, 30000, False),
                                                    (27, 256, 256, 700, True), (27, 32, 32, 5000, True), (8, 128, 96, 3000, False)])
def test_dgrad_epilogue_takes_batchnorm_reduction(kvol, c_in, c_out, n, relu):
    """b2m_conv_dgrad_bn_reduce: the dgrad whose result completes the gradient of a BatchNorm layer's output also
    accumulates that layer's backward reduction (sum g, sum g * xhat), g gated by the layer's ReLU bit mask - on the
    unsplit path (epilogue of conv_fwd_kernel) and on the offset-split path (conv_finalize_kernel). Must equal what
    b2m_bn_backward_reduce computes from the stored bf16 gradient, and the gradient itself must not change."""
    nbr_np, x, w, n = _conv_case(kvol, c_in, c_out, n, seed=21)
    torch.manual_seed(2)
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV)) if nbr_np is not None else None
    xd, wp = x.to(DEV).to(torch.bfloat16), ops.pack_weights(w.to(DEV), 0)
    pending = torch.randn(n, c_out, device=DEV).to(torch.bfloat16)
    px = (torch.randn(n, c_out, device=DEV) * 2 + 0.3).to(torch.bfloat16)          # the producer's pre-BatchNorm rows
    mean, invstd = torch.randn(c_out, device=DEV) * 0.2, torch.rand(c_out, device=DEV) + 0.5
    mask = torch.randint(0, 256, (n, c_out // 8), dtype=torch.uint8, device=DEV) if relu else None
    g0 = ops.conv_forward(xd, nbr, wp, kvol, n, c_out, residual=pending)
    red = torch.zeros(2 * c_out, dtype=torch.float64, device=DEV)
    g1 = ops.conv_forward(xd, nbr, wp, kvol, n, c_out, residual=pending, bn_reduce=(px, mask, mean, invstd, red))
    assert torch.equal(g0, g1)
    ref = torch.zeros(2 * c_out, dtype=torch.float64, device=DEV)
    from box2mask_b200 import _lib
    _lib.check(_lib.load().b2m_bn_backward_reduce(_lib.ptr(px), None, _lib.ptr(g1), n, c_out, _lib.ptr(mean), _lib.ptr(invstd),
                                                  1 if relu else 0, _lib.ptr(ref), _lib.ptr(mask), _lib.stream_ptr()), "reduce")
    # an independent fp64 evaluation as well
    gate = torch.ones(n, c_out, device=DEV, dtype=torch.float64)
    if relu:
        bits = (mask[:, :, None] >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1
        gate = bits.reshape(n, c_out).double()
    g = g1.double() * gate
    xhat = (px.double() - mean.double()) * invstd.double()
    exact = torch.cat([g.sum(0), (g * xhat).sum(0)])
    scale = exact.abs().max().item() + 1.0
    assert float((red - exact).abs().max()) <= 2e-5 * scale, float((red - exact).abs().max())
    assert float((red - ref).abs().max()) <= 4e-5 * scale


@pytest.mark.parametrize("kvol,c_in,c_out,n", [(27, 96, 96, 515), (27, 128, 96, 40000), (1, 96, 96, 30000), (8, 256, 256, 1500)])
def test_conv_forward_fused_epilogue(kvol, c_in, c_out, n):
    """Fused epilogue: v = acc * scale + shift + residual, ReLU (eval-mode BatchNorm folded into the convolution, the
    residual add of a BasicBlock), statistics of v, on the split and the unsplit path; fp32 output of the first c columns."""
    nbr_np, x, w, n = _conv_case(kvol, c_in, c_out, n, seed=11)
    torch.manual_seed(1)
    scale, shift = torch.rand(c_out) + 0.5, torch.randn(c_out)
    res = so.bf16_round(torch.randn(n, c_out))
    acc = so.sparse_conv(x.double(), nbr_np, w.double(), n_out=n)
    ref = torch.relu(acc * scale.double() + shift.double() + res.double())
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV)) if nbr_np is not None else None
    xd, wp = x.to(DEV).to(torch.bfloat16), ops.pack_weights(w.to(DEV), 0)
    cs = torch.zeros(2 * c_out, dtype=torch.float64, device=DEV)
    y = ops.conv_forward(xd, nbr, wp, kvol, n, c_out, cs, scale=scale.to(DEV), shift=shift.to(DEV),
                         residual=res.to(DEV).to(torch.bfloat16), relu=True)
    err = (y.float().cpu().double() - ref).abs()
    assert bool((err <= 8e-3 * ref.abs() + 2e-2).all()), float(err.max())
    assert torch.allclose(cs[:c_out].cpu(), ref.sum(0), rtol=1e-4, atol=2e-2)
    assert torch.allclose(cs[c_out:].cpu(), (ref * ref).sum(0), rtol=1e-4, atol=2e-2)
    # bias only (scale = None), no ReLU, fp32 logits of the first 13 columns
    y32 = ops.conv_forward(xd, nbr, wp, kvol, n, c_out, shift=shift.to(DEV), out_fp32_cols=13)
    ref32 = (acc + shift.double())[:, :13]
    assert y32.shape == (n, 13) and y32.dtype == torch.float32
    assert torch.allclose(y32.cpu().double(), ref32, rtol=1e-4, atol=1e-3)


def test_conv_forward_real_maps_and_dgrad(scene_coords):
    """Real kernel maps (k3 at L0, k2s2 down/up); dgrad through the mirrored / transposed weight packing."""
    c = torch.from_numpy(scene_coords).to(DEV)
    nbr3 = so.kernel_map_submanifold(scene_coords, 1, 3)
    coarse, parent = so.downsample_coords(scene_coords, 2)
    nbr_down, nbr_up = so.kernel_map_stride2(scene_coords, parent, len(coarse), 1)
    n = len(scene_coords)
    torch.manual_seed(0)
    x = so.bf16_round(torch.randn(n, 64)).double().requires_grad_(True)
    w = so.bf16_round(torch.randn(27, 64, 96) / 30).double().requires_grad_(True)
    y = so.sparse_conv(x, nbr3, w)
    dy = so.bf16_round(torch.randn(n, 96)).double()
    y.backward(dy)
    dev_nbr = ops.sort_kernel_map(torch.from_numpy(nbr3).to(DEV))
    got = ops.conv_forward(x.detach().float().to(DEV).to(torch.bfloat16), dev_nbr, ops.pack_weights(w.detach().float().to(DEV), 0), 27, n, 96)
    assert bool(((got.float().cpu().double() - y.detach()).abs() <= 8e-3 * y.detach().abs() + 2e-2).all())
    dx = ops.conv_forward(dy.float().to(DEV).to(torch.bfloat16), dev_nbr, ops.pack_weights(w.detach().float().to(DEV), 1), 27, n, 64)
    assert bool(((dx.float().cpu().double() - x.grad).abs() <= 8e-3 * x.grad.abs() + 2e-2).all())
    dw = ops.conv_wgrad(x.detach().float().to(DEV).to(torch.bfloat16), dy.float().to(DEV).to(torch.bfloat16), dev_nbr, 27, n)
    assert torch.allclose(dw.cpu().double(), w.grad, rtol=2e-3, atol=2e-3 * float(w.grad.abs().max()))
    # strided conv and its dgrad (mode 2 on nbr_up), transposed conv (mode 0 on nbr_up)
    xs = so.bf16_round(torch.randn(n, 32)).double().requires_grad_(True)
    ws = so.bf16_round(torch.randn(8, 32, 64) / 10).double().requires_grad_(True)
    ys = so.sparse_conv(xs, nbr_down, ws)
    dys = so.bf16_round(torch.randn(len(coarse), 64)).double()
    ys.backward(dys)
    d_down, d_up = ops.sort_kernel_map(torch.from_numpy(nbr_down).to(DEV)), ops.sort_kernel_map(torch.from_numpy(nbr_up).to(DEV))
    got = ops.conv_forward(xs.detach().float().to(DEV).to(torch.bfloat16), d_down, ops.pack_weights(ws.detach().float().to(DEV), 0), 8, len(coarse), 64)
    assert bool(((got.float().cpu().double() - ys.detach()).abs() <= 8e-3 * ys.detach().abs() + 2e-2).all())
    dxs = ops.conv_forward(dys.float().to(DEV).to(torch.bfloat16), d_up, ops.pack_weights(ws.detach().float().to(DEV), 2), 8, n, 32)
    assert bool(((dxs.float().cpu().double() - xs.grad).abs() <= 8e-3 * xs.grad.abs() + 2e-2).all())
    dws = ops.conv_wgrad(xs.detach().float().to(DEV).to(torch.bfloat16), dys.float().to(DEV).to(torch.bfloat16), d_down, 8, len(coarse))
    assert torch.allclose(dws.cpu().double(), ws.grad, rtol=2e-3, atol=2e-3 * float(ws.grad.abs().max()))


@pytest.mark.parametrize("kvol,c_in,c_out,n", [
    (27, 64, 64, 3000), (27, 128, 96, 6000), (27, 96, 96, 515), (27, 32, 32, 4000), (8, 256, 256, 2000),
    (27, 512, 256, 700), (125, 16, 32, 3000), (1, 64, 64, 1000), (8, 96, 128, 70000), (27, 192, 128, 63),
    (27, 48, 16, 400), (8, 32, 96, 5000), (27, 16, 16, 900), (1, 128, 96, 20000), (125, 8, 32, 3000), (27, 8, 96, 777),
])
def test_conv_wgrad(kvol, c_in, c_out, n):
    nbr_np, x, _, n = _conv_case(kvol, c_in, c_out, n, seed=kvol + c_out)
    dy = so.bf16_round(torch.randn(n, c_out))
    ref = torch.zeros(kvol, c_in, c_out, dtype=torch.float64)
    if nbr_np is None:
        ref[0] = x.double().t() @ dy.double()
    else:
        for k, (i, o) in enumerate(so.map_to_pairs(nbr_np)):
            if len(i):
                ref[k] = x.double()[i].t() @ dy.double()[o]
    nbr = ops.sort_kernel_map(torch.from_numpy(nbr_np).to(DEV)) if nbr_np is not None else None
    dw = ops.conv_wgrad(x.to(DEV).to(torch.bfloat16), dy.to(DEV).to(torch.bfloat16), nbr, kvol, n)
    # fp32 accumulation over up to 7e4 products + fp32 atomics across CTAs: rel 2e-3 of the largest entry
    assert torch.allclose(dw.cpu().double(), ref, rtol=2e-3, atol=2e-3 * float(ref.abs().max()) + 1e-4)


# ------------------------------------------------------------------------------------------------
# BatchNorm (+ residual + ReLU) against torch.nn.functional.batch_norm in fp32 on the same bf16 inputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,c,relu,res", [(5000, 96, True, True), (3001, 32, False, False), (700, 256, True, False),
                                          (2, 512, False, True), (20000, 128, True, True)])
def test_bn_forward_backward(n, c, relu, res):
    torch.manual_seed(n + c)
    x = so.bf16_round(torch.randn(n, c) * 2 + 0.5)
    r = so.bf16_round(torch.randn(n, c)) if res else None
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c)
    rm, rv = torch.zeros(c), torch.ones(c)
    xr = x.clone().requires_grad_(True)
    rr = r.clone().requires_grad_(True) if res else None
    g_ = gamma.clone().requires_grad_(True)
    b_ = beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = torch.nn.functional.batch_norm(xr, rm_ref, rv_ref, g_, b_, True, 0.1, 1e-5)
    if res:
        ref = ref + rr
    if relu:
        ref = torch.relu(ref)
    xd = x.to(DEV).to(torch.bfloat16)
    rd = r.to(DEV).to(torch.bfloat16) if res else None
    rm_d, rv_d = rm.to(DEV), rv.to(DEV)
    sums = ops.colstats(xd)
    assert torch.allclose(sums[:c].cpu(), x.double().sum(0), rtol=1e-5, atol=1e-2)
    out, sm, si = ops.bn_forward(xd, sums, gamma.to(DEV), beta.to(DEV), rm_d, rv_d, 0.1, 1e-5, True, rd, relu)
    # bf16 output rounding: 2^-8 relative + small absolute
    assert torch.allclose(out.float().cpu(), ref.detach(), rtol=8e-3, atol=1e-2)
    assert torch.allclose(rm_d.cpu(), rm_ref, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rv_d.cpu(), rv_ref, rtol=1e-4, atol=1e-5)
    dout = so.bf16_round(torch.randn(n, c))
    # the reference masks with ITS OWN relu output; use the device output so that near-zero ties agree
    ref_out_mask = (out.float().cpu() > 0) if relu else None
    ref2 = torch.nn.functional.batch_norm(xr, None, None, g_, b_, True, 0.1, 1e-5)
    if res:
        ref2 = ref2 + rr
    g = dout * ref_out_mask if relu else dout
    ref2.backward(g)
    dx, dres, dgamma, dbeta = ops.bn_backward(xd, out, dout.to(DEV).to(torch.bfloat16), sm, si, gamma.to(DEV), relu, True, res)
    assert torch.allclose(dx.float().cpu(), xr.grad, rtol=1e-2, atol=1e-2 * float(xr.grad.abs().max()) + 1e-3)
    assert torch.allclose(dgamma.cpu(), g_.grad, rtol=1e-3, atol=1e-3 * float(g_.grad.abs().max()) + 1e-2)
    assert torch.allclose(dbeta.cpu(), b_.grad, rtol=1e-3, atol=1e-3 * float(b_.grad.abs().max()) + 1e-2)
    if res:
        assert torch.allclose(dres.float().cpu(), rr.grad, rtol=8e-3, atol=1e-3)


def test_bn_backward_with_correlated_gradient():
    """Upstream gradients with a non-zero mean and a component along x-hat, so that the batch-statistics terms of
    the BatchNorm backward matter (they vanish for white-noise gradients)."""
    torch.manual_seed(1)
    n, c = 30000, 96
    x = so.bf16_round(torch.randn(n, c) * 1.5 + 0.3)
    gamma, beta = torch.rand(c) + 0.5, torch.randn(c)
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.functional.batch_norm(xr, None, None, gamma, beta, True, 0.1, 1e-5)
    xhat = ((x - x.mean(0)) / x.std(0, unbiased=False))
    dout = so.bf16_round(0.5 + 0.7 * xhat + 0.2 * torch.randn(n, c))
    ref.backward(dout)
    xd = x.to(DEV).to(torch.bfloat16)
    out, sm, si = ops.bn_forward(xd, ops.colstats(xd), gamma.to(DEV), beta.to(DEV), torch.zeros(c, device=DEV),
                                 torch.ones(c, device=DEV), 0.1, 1e-5, True, None, False)
    dx, _, dgamma, dbeta = ops.bn_backward(xd, out, dout.to(DEV).to(torch.bfloat16), sm, si, gamma.to(DEV), False, True, False)
    # dx is what is left after removing the mean and x-hat components: small, and it must still be right
    err = (dx.float().cpu() - xr.grad).abs()
    assert float(err.max()) <= 8e-3 * float(xr.grad.abs().max()) + 2e-3
    cos = torch.nn.functional.cosine_similarity(dx.float().cpu().flatten(), xr.grad.flatten(), dim=0)
    assert float(cos) > 0.999
    assert torch.allclose(dbeta.cpu(), dout.sum(0), rtol=1e-4) and torch.allclose(dgamma.cpu(), (dout * xhat).sum(0), rtol=2e-3, atol=1.0)


def test_bn_eval_mode():
    torch.manual_seed(0)
    n, c = 1000, 64
    x = so.bf16_round(torch.randn(n, c))
    gamma, beta, rm, rv = torch.rand(c) + 0.5, torch.randn(c), torch.randn(c) * 0.1, torch.rand(c) + 0.5
    ref = torch.nn.functional.batch_norm(x, rm, rv, gamma, beta, False, 0.1, 1e-5)
    out, _, _ = ops.bn_forward(x.to(DEV).to(torch.bfloat16), None, gamma.to(DEV), beta.to(DEV), rm.to(DEV), rv.to(DEV),
                               0.1, 1e-5, False, None, False)
    assert torch.allclose(out.float().cpu(), ref, rtol=8e-3, atol=1e-2)


# ------------------------------------------------------------------------------------------------
# pooling
# ------------------------------------------------------------------------------------------------
def test_segment_mean_max():
    torch.manual_seed(0)
    n, c, s = 20000, 96, 180
    ids = torch.randint(0, s - 2, (n,))
    ids[0] = s - 2                      # a segment of size 1 ...
    ids[1:4000] = s - 1                 # ... and a large one
    f = so.bf16_round(torch.randn(n, c))
    fd, idd = f.to(DEV).to(torch.bfloat16), ids.to(DEV)
    out, counts = ops.segment_mean_forward(fd, idd, s)
    ref = so.segment_mean(f.double(), ids, s)
    assert torch.allclose(out.cpu().double(), ref, rtol=1e-4, atol=1e-5)
    assert torch.equal(counts.cpu().long(), torch.bincount(ids, minlength=s))
    dout = torch.randn(s, c)
    df = ops.segment_mean_backward(dout.to(DEV), idd, counts, n)
    refd = dout[ids] / torch.bincount(ids, minlength=s)[ids][:, None]
    assert torch.allclose(df.float().cpu(), refd, rtol=8e-3, atol=1e-4)
    mx, arg = ops.segment_max_forward(fd, idd, s)
    assert torch.equal(mx.cpu(), so.segment_max(f, ids, s))
    assert torch.equal(f[arg.cpu().long(), torch.arange(c)[None].expand(s, c)], mx.cpu())


# ------------------------------------------------------------------------------------------------
# NMS / decode: bit-exact against the reference-generated golden vectors and the oracle
# ------------------------------------------------------------------------------------------------
def _check_nms(boxes, th, reps_ref, heat_ref, cluster_of_ref):
    reps, cluster_of, heat = ops.aabb_nms(boxes.to(DEV), th)
    assert np.array_equal(reps.cpu().numpy(), reps_ref)
    assert np.array_equal(cluster_of.cpu().numpy(), cluster_of_ref)
    assert np.array_equal(heat.cpu().numpy(), heat_ref)          # bit-exact fp32


@pytest.mark.parametrize("name", ["nms_small", "nms_medium", "nms_lowth"])
def test_nms_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    _check_nms(torch.from_numpy(g["boxes"]), float(g["th"]), g["reps"], g["heat"], g["cluster_of"])


def test_nms_hand_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_cases.npz"))
    for case in ["identical", "nested_eighth", "disjoint", "zero_volume", "chain"]:
        _check_nms(torch.from_numpy(g[case + "_boxes"]), 0.5, g[case + "_reps"], g[case + "_heat"], g[case + "_cluster_of"])


def test_nms_random_2k_vs_oracle():
    boxes = make_boxes(2000, 150, seed=7)
    reps, clusters, heat = onms.nms_clustering(boxes, 0.5)
    cof = np.full(len(boxes), -1, np.int32)
    for c, mem in enumerate(clusters):
        cof[mem.numpy()] = c
    _check_nms(boxes, 0.5, reps.numpy(), heat.numpy(), cof)
    # score ties resolve to the lower index (stable order)
    tb = boxes[:200].clone()
    tb[:, 0] = torch.round(tb[:, 0] * 4) / 4
    reps, clusters, heat = onms.nms_clustering(tb, 0.5)
    cof = np.full(len(tb), -1, np.int32)
    for c, mem in enumerate(clusters):
        cof[mem.numpy()] = c
    _check_nms(tb, 0.5, reps.numpy(), heat.numpy(), cof)


def test_heatmap_project_and_mask_nms(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms_medium.npz"))
    heat = torch.from_numpy(g["heat"])                 # [K, M_fg]
    k, m_fg = heat.shape
    rng = np.random.default_rng(0)
    s = m_fg + 57
    fg_mask = np.zeros(s, bool)
    fg_mask[rng.permutation(s)[:m_fg]] = True
    n_vox = 10007
    seg2vox = torch.from_numpy(rng.integers(0, s, n_vox))
    full = torch.zeros(k, s)
    full[:, torch.from_numpy(fg_mask)] = heat
    ref_masks = full[:, seg2vox] > 0.3
    fg_rank = torch.full((s,), -1, dtype=torch.int32)
    fg_rank[torch.from_numpy(fg_mask)] = torch.arange(m_fg, dtype=torch.int32)
    packed = ops.heatmap_project(heat.to(DEV), fg_rank.to(DEV), seg2vox.to(DEV), 0.3)
    assert torch.equal(ops.unpack_masks(packed, n_vox).cpu(), ref_masks)
    assert torch.equal(ops.pack_masks_torch(ref_masks.to(DEV)), packed)
    nonempty = ref_masks.sum(1) > 0
    keep_ref = onms.mask_nms(ref_masks[nonempty], 0.6)
    keep = ops.mask_nms(ops.pack_masks_torch(ref_masks[nonempty].to(DEV)), 0.6)
    assert np.array_equal(torch.nonzero(keep).flatten().cpu().numpy(), keep_ref.numpy())
    # the fixture's own mask-NMS result (masks = heat > 0.3 in fg space)
    keep2 = ops.mask_nms(ops.pack_masks_torch((heat > 0.3).to(DEV)), 0.6)
    assert np.array_equal(torch.nonzero(keep2).flatten().cpu().numpy(), g["mask_keep"])


def test_detection2mask_matches_reference_golden(golden_dir):
    """The GPU decode (box2mask_b200.decode.detection2mask: NMS clustering, heat-map projection, mask-NMS, label vote,
    point projection) against the reference's own detection2mask run on the CPU (tests/golden/decode_small.npz)."""
    from types import SimpleNamespace
    from box2mask_b200 import decode
    from box2mask_b200.selection_net import default_config
    from box2mask_b200.synthetic import label_maps
    from oracle.make_golden import decode_inputs
    g = np.load(os.path.join(golden_dir, "decode_small.npz"))
    cfg = default_config()
    valid, _, is_fg = label_maps(20)
    batch, pred = decode_inputs()
    net = SimpleNamespace(device=DEV, semantic_valid_class_ids=valid, is_foreground=is_fg)
    for mode in ("eval", "train"):
        res = decode.detection2mask(net, batch, pred, cfg, mode, True, *cfg.eval_ths)
        for scene in batch["scene"]:
            name, r = scene["name"], res[scene["name"]]
            # scores go through sigmoid on the device: allow the last ulp, everything discrete must be identical
            assert np.allclose(r["conf"].numpy(), g["%s_%s_conf" % (mode, name)], rtol=0, atol=2e-7)
            assert np.array_equal(np.asarray(r["label_id"]), g["%s_%s_label_id" % (mode, name)])
            shape = tuple(g["%s_%s_mask_shape" % (mode, name)])
            ref = np.unpackbits(g["%s_%s_mask" % (mode, name)], axis=1)[:, :shape[1]].astype(bool)
            assert tuple(r["mask"].shape) == shape
            assert np.array_equal(r["mask"].numpy(), ref)
            if mode == "train":
                assert np.array_equal(r["cluster_representatives"].numpy(), g["train_%s_reps" % name])


def test_detection2mask_s3dis_branch_matches_reference_golden(golden_dir):
    """S3DIS decode (per-voxel semantics head -> per-segment mode, no mask-NMS; models/detection_net.py:378-381,395-415,
    449-451) on the GPU against the reference's own detection2mask (tests/golden/decode_s3dis.npz)."""
    from types import SimpleNamespace
    from box2mask_b200 import decode
    from box2mask_b200.synthetic import label_maps
    from oracle.make_golden import decode_inputs_s3dis, variant_config
    g = np.load(os.path.join(golden_dir, "decode_s3dis.npz"))
    cfg, _, _ = variant_config("s3dis")
    valid, _, is_fg = label_maps(13)
    batch, pred = decode_inputs_s3dis()
    net = SimpleNamespace(device=DEV, semantic_valid_class_ids=valid, is_foreground=is_fg, requires_voxel_outputs=True)
    name = batch["scene"][0]["name"]
    for mode in ("eval", "train"):
        r = decode.detection2mask(net, batch, pred, cfg, mode, True, *cfg.eval_ths)[name]
        assert np.allclose(r["conf"].numpy(), g["%s_%s_conf" % (mode, name)], rtol=0, atol=2e-7)
        assert np.array_equal(np.asarray(r["label_id"]), g["%s_%s_label_id" % (mode, name)])
        shape = tuple(g["%s_%s_mask_shape" % (mode, name)])
        ref = np.unpackbits(g["%s_%s_mask" % (mode, name)], axis=1)[:, :shape[1]].astype(bool)
        assert tuple(r["mask"].shape) == shape and np.array_equal(r["mask"].numpy(), ref)
        if mode == "train":
            assert np.array_equal(r["cluster_representatives"].numpy(), g["train_%s_reps" % name])


def test_label_vote_kernels():
    rng = np.random.default_rng(3)
    n_vox, k, n_lab = 20011, 37, 21
    label = rng.integers(0, n_lab, n_vox).astype(np.int32)
    masks = rng.random((k, n_vox)) < rng.random((k, 1)) * 0.3
    masks[5] = False                                                     # an empty mask votes for label 0
    packed = ops.pack_masks_torch(torch.from_numpy(masks).to(DEV))
    best, counts = ops.mask_label_vote(packed, torch.from_numpy(label).to(DEV), n_vox, n_lab)
    ref_counts = np.stack([np.bincount(label[m], minlength=n_lab) for m in masks])
    assert np.array_equal(counts.cpu().numpy(), ref_counts)
    assert np.array_equal(best.cpu().numpy(), ref_counts.argmax(1))       # np.argmax: lowest label on ties
    seg = rng.integers(0, 500, n_vox)
    seg[:40] = 7
    label[:40] = np.tile([4, 2], 20)                                      # a tie: torch.mode / np.argmax take the lower label
    label[seg == 7] = np.where(np.arange((seg == 7).sum()) % 2 == 0, 4, 2)
    best, counts = ops.segment_label_vote(torch.from_numpy(seg).to(DEV), torch.from_numpy(label).to(DEV), 500, n_lab)
    for sid in (0, 7, 123, 499):
        sel = torch.from_numpy(label[seg == sid]).long()
        if len(sel):
            assert int(best[sid]) == int(torch.mode(sel)[0]), sid
    assert int(counts.sum()) == n_vox


def test_nms_clustering_drop_in_cluster_lists(golden_dir):
    """decode.NMS_clustering keeps the reference's return values: representatives, per-cluster member lists in
    descending-score order, heat-maps (models/iou_nms.py:68-105) — against the reference-generated fixture."""
    from box2mask_b200.decode import NMS_clustering
    g = np.load(os.path.join(golden_dir, "nms_medium.npz"))
    reps, clusters, heat = NMS_clustering(torch.from_numpy(g["boxes"]).to(DEV), float(g["th"]))
    assert np.array_equal(reps.cpu().numpy(), g["reps"]) and np.array_equal(heat.cpu().numpy(), g["heat"])
    assert [len(c) for c in clusters] == g["cluster_sizes"].tolist()
    assert np.array_equal(torch.cat(clusters).cpu().numpy(), g["cluster_members"])


def test_sparse_tensor_validates_and_deduplicates():
    """ME.SparseTensor contract (models/model.py:43): out-of-range coordinates raise, duplicate coordinates collapse onto
    their first occurrence with the features re-indexed (unique_index / inverse_mapping); unique input keeps its order."""
    import box2mask_b200
    ME = box2mask_b200.install_as_minkowski_engine()
    rng = np.random.default_rng(0)
    base = np.unique(rng.integers(0, 40, (3000, 3)), axis=0).astype(np.int32)
    coords = np.concatenate([np.zeros((len(base), 1), np.int32), base], 1)
    feats = torch.from_numpy(rng.normal(size=(len(coords), 6)).astype(np.float32))
    st = ME.SparseTensor(feats, torch.from_numpy(coords), device=DEV)
    assert st.unique_index is None and torch.equal(st.C.cpu(), torch.from_numpy(coords)) and torch.equal(st.F.cpu(), feats)
    dup_rows = rng.integers(0, len(coords), 500)
    c2 = np.concatenate([coords, coords[dup_rows]], 0)
    perm = rng.permutation(len(c2))
    c2 = c2[perm]
    f2 = torch.from_numpy(rng.normal(size=(len(c2), 6)).astype(np.float32))
    st2 = ME.SparseTensor(f2, torch.from_numpy(c2), device=DEV)
    assert len(st2) == len(coords)
    _, first = np.unique(c2, axis=0, return_index=True)
    ui = np.sort(first)                                                   # first occurrences, in input order
    assert np.array_equal(st2.unique_index.cpu().numpy(), ui)
    assert torch.equal(st2.C.cpu(), torch.from_numpy(c2[ui])) and torch.equal(st2.F.cpu(), f2[ui])
    assert np.array_equal(c2[ui][st2.inverse_mapping.cpu().numpy()], c2)
    bad = coords.copy()
    bad[3, 2] = 40000
    with pytest.raises(RuntimeError):
        ME.SparseTensor(feats, torch.from_numpy(bad), device=DEV)


@pytest.mark.parametrize("n,c,res", [(5000, 96, True), (70001, 256, False), (3, 32, True)])
def test_bn_relu_mask_equals_output_gate(n, c, res):
    """The 1-bit ReLU gate written by bn_forward (uint8[n, c/8]) against the gate re-read from `out`: the backward
    passes must give identical bits either way (dx, dresidual, dgamma, dbeta)."""
    torch.manual_seed(n + c)
    x = torch.randn(n, c, device=DEV).to(torch.bfloat16)
    r = torch.randn(n, c, device=DEV).to(torch.bfloat16) if res else None
    gamma, beta = torch.rand(c, device=DEV) + 0.5, torch.randn(c, device=DEV) * 0.1
    out, sm, si, mask = ops.bn_forward(x, ops.colstats(x), gamma, beta, torch.zeros(c, device=DEV), torch.ones(c, device=DEV),
                                       0.1, 1e-5, True, r, True, want_mask=True)
    bits = (mask[:, :, None] >> torch.arange(8, device=DEV, dtype=torch.uint8)) & 1
    assert torch.equal(bits.reshape(n, c).bool(), out > 0)
    dout = torch.randn(n, c, device=DEV).to(torch.bfloat16)
    a = ops.bn_backward(x, out, dout, sm, si, gamma, True, True, res)
    b = ops.bn_backward(x, None, dout, sm, si, gamma, True, True, res, relu_mask=mask)
    for u, v in zip(a, b):
        assert (u is None and v is None) or torch.equal(u, v)
