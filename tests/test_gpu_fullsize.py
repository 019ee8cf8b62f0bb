"""Full-size checks (BASELINE.json sizes: a batch of ScanNet-shape scenes, ~150k voxels each, and one S3DIS-shape
room, ~0.75M voxels) through size-independent properties that have EXACT answers, so no CPU oracle has to run at
that size:

* kernel maps: the submanifold map is an involution (nbr[k][o] == i  <=>  nbr[K-1-k][i] == o), its centre offset is the
  identity, the stride-2 map places every fine row exactly once and the transposed map is its inverse;
* convolution forward / dgrad / wgrad with small-integer features and signed-permutation weights: every partial sum
  is an integer below 2^8 (forward, exactly representable in bf16) or 2^24 (wgrad, exact in fp32 whatever the
  summation order), so the CUDA result must equal a plain torch index computation BIT FOR BIT;
* superpoint mean pooling of a per-segment constant returns that constant exactly.
All calls go through the C-ABI (box2mask_b200.ops)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from box2mask_b200 import ops  # noqa: E402
from box2mask_b200.synthetic import batched_coordinates, make_scene  # noqa: E402

DEV = "cuda"


def _level0(kind):
    if kind == "scannet_batch":
        scenes = [make_scene(20000 + i, scale=0.84) for i in range(3)]       # 3 x ~150k voxels
    else:
        scenes = [make_scene(30000, scale=2.18)]                             # one S3DIS-shape room, ~0.75M voxels
    coords = batched_coordinates([s["vox_coords"] for s in scenes]).to(DEV)
    segs = [torch.from_numpy(s["vox_segments"]) for s in scenes]
    return coords, segs


@pytest.fixture(scope="module", params=["scannet_batch", "s3dis_room"])
def level0(request):
    return _level0(request.param)


def _signed_permutations(kvol, c_in, c_out, seed):
    """W[k] = a signed partial permutation matrix [c_in, c_out]: exact in bf16, one non-zero per output column."""
    g = torch.Generator().manual_seed(seed)
    w = torch.zeros(kvol, c_in, c_out)
    for k in range(kvol):
        src = torch.randint(0, c_in, (c_out,), generator=g)
        sign = torch.randint(0, 2, (c_out,), generator=g).float() * 2 - 1
        w[k, src, torch.arange(c_out)] = sign
    return w


def _reference_conv(x, nbr, w, n_out):
    """y[o] = sum_k x[nbr[k][o]] @ w[k] with fp32 index ops on the device (exact for small integers)."""
    xz = torch.cat([x.float(), torch.zeros(1, x.shape[1], device=x.device)], 0)     # row -1 -> zeros
    y = torch.zeros(n_out, w.shape[2], device=x.device)
    for k in range(w.shape[0]):
        y += xz[nbr[k, :n_out].long()] @ w[k].to(x.device)
    return y


def test_kernel_map_properties_full_size(level0):
    coords, _ = level0
    n = coords.shape[0]
    table = ops.hash_build(coords)
    assert table.status.cpu().tolist() == [0, 0]
    nbr = ops.kernel_map_submanifold(coords, 1, 3, table)
    assert torch.equal(nbr[13, :n].cpu(), torch.arange(n, dtype=torch.int32))            # centre offset = identity
    assert bool((nbr[:, n:] == -1).all())                                                  # padding
    rows = torch.arange(n, dtype=torch.int32, device=DEV)
    for k in range(13):                                                                    # involution, all 13 pairs
        fwd, back = nbr[k, :n], nbr[26 - k, :n]
        has = fwd >= 0
        assert torch.equal(back[fwd[has].long()], rows[has])
        assert int(has.sum()) == int((back >= 0).sum())
    coarse, parent = ops.downsample_coords(coords, 2)
    m = coarse.shape[0]
    # strided coordinates = floor to a multiple of 2, sorted unique
    assert torch.equal(coarse[parent.long()][:, 1:], coords[:, 1:] - (coords[:, 1:] % 2))
    assert torch.equal(coarse[parent.long()][:, 0], coords[:, 0])
    key = ((coarse[:, 0].long() * 4096 + coarse[:, 1]) * 4096 + coarse[:, 2]) * 4096 + coarse[:, 3]
    assert bool((key[1:] > key[:-1]).all())
    down, up = ops.kernel_map_stride2(coords, parent, m, 1)
    placed = down[:, :m][down[:, :m] >= 0]
    assert placed.numel() == n and torch.equal(torch.sort(placed).values, rows)           # every fine row exactly once
    for k in range(8):
        has = up[k, :n] >= 0
        assert torch.equal(down[k, :m][up[k, :n][has].long()], rows[has])                 # transposed map = inverse
    assert int((up[:, :n] >= 0).sum()) == n


@pytest.mark.parametrize("c_in,c_out", [(96, 96), (128, 96), (32, 64)])
def test_conv_exact_integers_full_size(level0, c_in, c_out):
    coords, _ = level0
    n = coords.shape[0]
    nbr = ops.kernel_map_submanifold(coords, 1, 3, ops.hash_build(coords))
    km = ops.sort_kernel_map(nbr, n)
    g = torch.Generator().manual_seed(c_in + c_out)
    x = torch.randint(-4, 5, (n, c_in), generator=g).float().to(DEV)
    w = _signed_permutations(27, c_in, c_out, seed=c_in)
    # forward: |y| <= 27 * 4 = 108 < 256, exactly representable in bf16
    y = ops.conv_forward(x.to(torch.bfloat16), km, ops.pack_weights(w.to(DEV), 0), 27, n, c_out)
    ref = _reference_conv(x, nbr, w, n)
    assert torch.equal(y.float(), ref)
    # dgrad (mirrored, transposed weights over the same map): dx[i] = sum_k dy[nbr[26-k][i]] @ w[k]^T
    dy = torch.randint(-4, 5, (n, c_out), generator=g).float().to(DEV)
    wt = w.to(DEV)
    dx = ops.conv_forward(dy.to(torch.bfloat16), km, ops.pack_weights(wt, 1), 27, n, c_in)
    # a signed partial permutation can map several output columns to one input channel: |dx| <= 27 * 4 * (fan-in)
    ref_dx = _reference_conv(dy, nbr, torch.flip(wt, [0]).transpose(1, 2).contiguous(), n)
    # the fp32 accumulation is exact (integers far below 2^24); beyond 256 the bf16 store rounds, identically on both sides
    assert torch.equal(dx.float(), ref_dx.to(torch.bfloat16).float())
    # wgrad: sums of products of small integers over up to ~2.3M pairs stay below 2^24 -> exact in fp32 in any order
    x2 = torch.randint(-2, 3, (n, c_in), generator=g).float().to(DEV)
    dy2 = torch.randint(-1, 2, (n, c_out), generator=g).float().to(DEV)
    dw = ops.conv_wgrad(x2.to(torch.bfloat16), dy2.to(torch.bfloat16), km, 27, n)
    xz = torch.cat([x2, torch.zeros(1, c_in, device=DEV)], 0)
    for k in (0, 5, 13, 26):
        ref_dw = xz[nbr[k, :n].long()].t() @ dy2          # fp32 matmul of integers below 2^24: exact
        assert torch.equal(dw[k], ref_dw)


def test_strided_and_transposed_conv_exact_full_size(level0):
    coords, _ = level0
    n = coords.shape[0]
    coarse, parent = ops.downsample_coords(coords, 2)
    m = coarse.shape[0]
    down, up = ops.kernel_map_stride2(coords, parent, m, 1)
    kd, ku = ops.sort_kernel_map(down, m), ops.sort_kernel_map(up, n)
    g = torch.Generator().manual_seed(5)
    w = _signed_permutations(8, 96, 96, seed=9)
    x = torch.randint(-4, 5, (n, 96), generator=g).float().to(DEV)
    y = ops.conv_forward(x.to(torch.bfloat16), kd, ops.pack_weights(w.to(DEV), 0), 8, m, 96)     # |y| <= 8 * 4
    assert torch.equal(y.float(), _reference_conv(x, down, w, m))
    z = torch.randint(-4, 5, (m, 96), generator=g).float().to(DEV)
    yt = ops.conv_forward(z.to(torch.bfloat16), ku, ops.pack_weights(w.to(DEV), 0), 8, n, 96)   # transposed conv
    assert torch.equal(yt.float(), _reference_conv(z, up, w, n))


def test_segment_mean_constant_full_size(level0):
    coords, segs = level0
    off, ids = 0, []
    for s in segs:
        ids.append(s + off)
        off += int(s.max()) + 1
    ids = torch.cat(ids).to(DEV)
    value = (torch.arange(off, device=DEV) % 251).float()                  # exact in bf16
    f = value[ids][:, None].repeat(1, 96).to(torch.bfloat16)
    pooled, counts = ops.segment_mean_forward(f, ids, off)
    assert torch.equal(pooled.float(), value[:, None].repeat(1, 96))
    assert int(counts.sum()) == coords.shape[0]
