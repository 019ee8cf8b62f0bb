"""CPU: the trunk executor's hand-scheduled forward / backward (box2mask_b200/trunk.py) against the module-by-module
autograd path, with the C-ABI ops replaced by plain torch stand-ins (float32 tensors in place of bf16). This checks the
SCHEDULE - unit order, residual / skip-connection gradient fusion into the dgrad epilogue, concatenation splits, the
flat gradient buffer, BatchNorm bookkeeping, eval-mode folding - independently of the CUDA kernels, which the GPU
tests cover. The stand-ins restate the ops' contracts (include/b2m.h); the kernel maps come from the oracle."""
import numpy as np
import pytest
import torch

from box2mask_b200 import ops
from box2mask_b200.me import nn as me_nn
from box2mask_b200.me.sparse_tensor import SparseTensor
from box2mask_b200.selection_net import SelectionNet, default_config
from box2mask_b200.synthetic import make_batch
from oracle import sparse_ops as so


class FakeMap:
    def __init__(self, nbr):
        self.nbr = torch.from_numpy(np.asarray(nbr)).long()
        self.kvol, self.n_out = self.nbr.shape
        self.order = self.gmask = None


class FakeManager:
    def __init__(self, coords):
        self.levels = {1: np.asarray(coords)}
        self.stride2, self.sub = {}, {}

    def submanifold_map(self, ts, k):
        if (ts, k) not in self.sub:
            self.sub[(ts, k)] = FakeMap(so.kernel_map_submanifold(self.levels[ts], ts, k))
        return self.sub[(ts, k)]

    def stride2_maps(self, ts):
        if ts not in self.stride2:
            coarse, parent = so.downsample_coords(self.levels[ts], 2 * ts)
            self.levels[2 * ts] = coarse
            down, up = so.kernel_map_stride2(self.levels[ts], parent, len(coarse), ts)
            self.stride2[ts] = (FakeMap(down), FakeMap(up))
        return self.stride2[ts]

    def coords(self, ts):
        return torch.from_numpy(self.levels[ts])

    def wait_ready(self):
        pass

    def prepare(self, n_strided, sub_kernels, stream=None):
        s = 1
        for _ in range(n_strided):
            self.stride2_maps(s)
            s *= 2
        for ts, k in sub_kernels:
            self.submanifold_map(ts, k)


def _gather_mm(x, nbr, w):
    """y[o] = sum_k x[nbr[k][o]] @ w[k] (entries < 0 contribute zero)"""
    n_out = nbr.shape[1] if nbr is not None else x.shape[0]
    y = torch.zeros(n_out, w.shape[-1], dtype=torch.float64)
    if nbr is None:
        return x.double() @ w[0].double()
    for k in range(w.shape[0]):
        sel = nbr[k] >= 0
        if bool(sel.any()):
            y[sel] += x.double()[nbr[k][sel]] @ w[k].double()
    return y


fused_reductions = []       # one entry per dgrad call that took a BatchNorm reduction (reset by the tests that count)


@pytest.fixture()
def fake_ops(monkeypatch):
    def cast_pad_bf16(x, c_pad):
        return torch.nn.functional.pad(x, (0, c_pad - x.shape[1]))

    def conv_forward(x, kmap, packed_w, kvol, n_out, c_n, colsum=None, scale=None, shift=None, residual=None, relu=False,
                     out_fp32_cols=None, bn_reduce=None):
        v = _gather_mm(x, kmap.nbr if kmap is not None else None, packed_w)
        assert v.shape == (n_out, c_n)
        if scale is not None:
            v = v * scale.double()
        if shift is not None:
            v = v + shift.double()
        if residual is not None:
            v = v + residual.double()
        if relu:
            v = torch.relu(v)
        if colsum is not None:
            colsum[:c_n] += v.sum(0)
            colsum[c_n:2 * c_n] += (v * v).sum(0)
        if bn_reduce is not None:       # b2m_conv_dgrad_bn_reduce: the producer layer's BatchNorm-backward reduction
            px, pmask, pmean, pinvstd, red = bn_reduce
            fused_reductions.append(1)
            g = v.float().double() * (pmask if pmask is not None else 1)
            red[:c_n] += g.sum(0)
            red[c_n:2 * c_n] += (g * (px.double() - pmean.double()) * pinvstd.double()).sum(0)
        return v.float()

    def conv_wgrad(x, dy, kmap, kvol, n_out, out=None):
        dw = torch.zeros(kvol, x.shape[1], dy.shape[1], dtype=torch.float64)
        if kmap is None:
            dw[0] = x.double().t() @ dy.double()
        else:
            for k in range(kvol):
                sel = kmap.nbr[k] >= 0
                if bool(sel.any()):
                    dw[k] = x.double()[kmap.nbr[k][sel]].t() @ dy.double()[sel]
        if out is not None:
            out.copy_(dw.float().reshape(out.shape))
            return out
        return dw.float()

    def pack(kernel, c_in, mode):
        w = kernel.detach()
        if w.dim() == 2:
            w = w[None]
        if w.shape[1] != c_in:
            w = torch.nn.functional.pad(w, (0, 0, 0, c_in - w.shape[1]))
        if mode == 0:
            return w
        wt = w.transpose(1, 2)
        return wt.flip(0) if mode == 1 else wt          # mode 1: mirrored offsets, mode 2: plain transpose

    def prepack(module):
        for m in module.modules():
            if isinstance(m, me_nn._ConvBase) and not (m.kernel_volume == 1 and m.stride == 1 and (
                    m.bias is not None or m.out_channels % 16 != 0)):
                c_in = me_nn._round16(m.in_channels)
                m.__dict__["_prepacked"] = (pack(m.kernel, c_in, 0), pack(m.kernel, c_in, me_nn._dgrad_mode(m)),
                                            m.kernel._version, c_in, me_nn._dgrad_mode(m))

    def pack_weights(kernel, mode):
        return pack(kernel, kernel.shape[-2], mode)

    def take(n, device):
        return torch.zeros(n, dtype=torch.float64)

    def colstats(x):
        return torch.cat([x.double().sum(0), (x.double() ** 2).sum(0)])

    def bn_forward(x, sums, gamma, beta, running_mean, running_var, momentum, eps, training, residual=None, relu=False,
                   n_stat=None, want_mask=False):
        n = x.shape[0] if n_stat is None else n_stat
        c = x.shape[1]
        if training:
            mean = sums[:c] / n
            var = (sums[c:2 * c] / n - mean * mean).clamp(min=0)
            running_mean.mul_(1 - momentum).add_(momentum * mean.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * n / max(n - 1, 1)).float())
        else:
            mean, var = running_mean.double(), running_var.double()
        invstd = 1.0 / torch.sqrt(var + eps)
        out = (x.double() - mean) * invstd * gamma.double() + beta.double()
        if residual is not None:
            out = out + residual.double()
        if relu:
            out = torch.relu(out)
        if want_mask:           # the stand-in keeps the gate as a boolean matrix (the kernels pack it into bits)
            return out.float(), mean.float(), invstd.float(), (out > 0) if relu else None
        return out.float(), mean.float(), invstd.float()

    def bn_backward(x, out, dout, save_mean, save_invstd, gamma, relu, training, want_dresidual, n_stat=None,
                    reduce_hook=None, n_stat_dev=None, dgamma=None, dbeta=None, relu_mask=None, red=None):
        n = x.shape[0] if n_stat is None else n_stat
        c = x.shape[1]
        g = dout.double()
        if relu:
            g = g * (relu_mask if relu_mask is not None else (out > 0))
        xhat = (x.double() - save_mean.double()) * save_invstd.double()
        sg, sgx = g.sum(0), (g * xhat).sum(0)
        if red is not None:             # the reduction a dgrad epilogue took must be the one this layer would compute
            assert torch.allclose(red[:c], sg, rtol=1e-9, atol=1e-9) and torch.allclose(red[c:2 * c], sgx, rtol=1e-9, atol=1e-9)
            sg, sgx = red[:c].clone(), red[c:2 * c].clone()
        scale = gamma.double() * save_invstd.double()
        dx = scale * (g - sg / n - xhat * sgx / n) if training else scale * g
        for dst, val in ((dgamma, sgx), (dbeta, sg)):
            if dst is not None:
                dst.copy_(val.float())
        return dx.float(), (g.float() if want_dresidual else None), (dgamma if dgamma is not None else sgx.float()), \
            (dbeta if dbeta is not None else sg.float())

    def segment_mean_forward(f, ids, s):
        out = torch.zeros(s, f.shape[1]).index_add_(0, ids, f.float())
        cnt = torch.bincount(ids, minlength=s).float()
        return out / cnt[:, None].clamp(min=1), cnt

    def segment_mean_backward(dout, ids, counts, n):
        return dout[ids] / counts[ids][:, None]

    for name, fn in dict(cast_pad_bf16=cast_pad_bf16, conv_forward=conv_forward, conv_wgrad=conv_wgrad, colstats=colstats,
                         bn_forward=bn_forward, bn_backward=bn_backward, pack_weights=pack_weights,
                         segment_mean_forward=segment_mean_forward, segment_mean_backward=segment_mean_backward).items():
        monkeypatch.setattr(ops, name, fn)
    monkeypatch.setattr(ops.ZeroArena, "take", classmethod(lambda cls, n, device: take(n, device)))
    monkeypatch.setattr(ops.ZeroArena, "generation", classmethod(lambda cls, device: 0))
    monkeypatch.setattr(me_nn, "prepack_conv_weights", prepack)
    import box2mask_b200.me as ME
    monkeypatch.setattr(ME, "prepack_conv_weights", prepack)
    return prepack


def _net_and_batch(seed=0):
    torch.manual_seed(seed)
    cfg = default_config()
    net = SelectionNet(cfg, "cpu", list(range(20)), out_channels=[96, 96, 6])
    with torch.no_grad():                          # non-trivial BatchNorm state
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    batch = _line_batch()
    return cfg, net, batch


def _line_batch(n_scenes=4, length_vox=2250, seed=0):
    """Thin lines of voxels along x (2 x 2 voxels across): ~6k voxels per scene but 18 cells at tensor stride 128, so the
    deepest level of the batch has 72 rows and training-mode BatchNorm is well conditioned on every level (a random
    room-shaped toy scene leaves 2-4 rows there, which makes the comparison chaotic)."""
    from box2mask_b200.synthetic import collate
    rng = np.random.default_rng(seed)
    scenes = []
    for b in range(n_scenes):
        keep = rng.random((length_vox, 2, 2)) < 0.8
        coords = np.stack(np.nonzero(keep), 1).astype(np.int32)
        n = len(coords)
        _, segs = np.unique(coords[:, 0] // 16, return_inverse=True)
        s = int(segs.max()) + 1
        scenes.append({"vox_coords": coords, "vox_features": rng.normal(0, 1, (n, 6)).astype(np.float32),
                       "vox_segments": segs.astype(np.int64), "input_location": rng.normal(0, 1, (s, 3)).astype(np.float32),
                       "gt_bb_offsets": np.zeros((s, 3), np.float32), "gt_bb_bounds": np.ones((s, 3), np.float32),
                       "gt_semantics": np.zeros(s, np.int64), "fg_instances": np.ones(s, bool)})
    return collate(scenes)


def _run(net, batch, executor):
    net.use_trunk_executor = executor
    cm = FakeManager(batch["vox_coords"].numpy())
    x = SparseTensor(batch["vox_features"], coordinate_manager=cm)
    out = net(x, batch["pooling_ids"])
    return out


def test_trunk_executor_train_matches_module_path(fake_ops):
    cfg, net, batch = _net_and_batch()
    net.train()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    res = {}
    net.trunk_executor().fuse_bn_reduce = "all"     # every fusable BatchNorm-backward reduction goes through a dgrad epilogue
    del fused_reductions[:]
    for executor in (False, True):
        net.load_state_dict(sd)
        net.zero_grad(set_to_none=True)
        out = _run(net, batch, executor)
        loss = sum((out[h].F.float() ** 2).mean() * (i + 1) for i, h in enumerate(cfg.network_heads))
        loss.backward()
        res[executor] = ({h: out[h].F.detach().clone() for h in cfg.network_heads},
                         {k: p.grad.detach().clone() for k, p in net.named_parameters()},
                         {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "tracked" in k})
    (o0, g0, b0), (o1, g1, b1) = res[False], res[True]
    # the module path normalises in fp32 (torch BatchNorm1d on the float32 stand-ins), the stand-in ops in fp64, and
    # training-mode BatchNorm over the 2-4 rows of the deepest levels amplifies that: 5e-3 on outputs, cosine on gradients
    for h in o0:
        assert torch.allclose(o0[h], o1[h], rtol=5e-3, atol=5e-3), (h, float((o0[h] - o1[h]).abs().max()))
    assert set(g0) == set(g1)
    worst = {}
    for k in g0:
        if float(g0[k].norm()) == 0:
            assert float(g1[k].norm()) == 0, k
            continue
        worst[k] = float(torch.nn.functional.cosine_similarity(g0[k].flatten().double(), g1[k].flatten().double(), dim=0))
        assert abs(float(g1[k].norm()) / float(g0[k].norm()) - 1) < 2e-2, (k, float(g0[k].norm()), float(g1[k].norm()))
    assert min(worst.values()) > 0.9995, sorted(worst.items(), key=lambda kv: kv[1])[:5]
    for k in b0:
        assert torch.allclose(b0[k].float(), b1[k].float(), rtol=1e-3, atol=1e-4), k
    # trunk gradients are slices of ONE flat buffer laid out in order of completion during backward
    ex = net.trunk_executor()
    flat = ex.grads.flat
    for p in ex.program.parameters():
        assert flat.data_ptr() <= p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4
    names = [u.name for u in ex.program.units]
    assert names[0] == "conv0p1s1" and names[-1] == "block8.1.conv2" and len(names) == 81
    assert ex.grads.index[id(net.block8[1].conv2.kernel)] == 0           # the last unit's gradients complete first
    # 63 of the 81 BatchNorm-backward reductions were taken by the dgrad that completes the layer's output gradient (the
    # stand-in bn_backward asserts each one equals the reduction the layer would have computed itself); not fusable:
    # the 7 transposed convolutions (concatenation split), the 10 residual-branch 1x1 units, the trunk output
    assert len(fused_reductions) == 63 == sum(u.fuse_for is not None for u in ex.program.units)


def test_trunk_executor_accumulates_when_gradients_are_not_zeroed(fake_ops):
    cfg, net, batch = _net_and_batch(seed=1)
    net.train()
    net.use_trunk_executor = True

    def step():
        out = _run(net, batch, True)
        sum((out[h].F.float() ** 2).mean() for h in cfg.network_heads).backward()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    step()
    g1 = net.block8[0].conv1.kernel.grad.clone()
    net.load_state_dict(sd)            # same BatchNorm running statistics -> the same gradients again
    step()                             # no zero_grad in between: gradients accumulate like autograd's
    assert torch.allclose(net.block8[0].conv1.kernel.grad, 2 * g1, rtol=1e-3, atol=1e-7)


def test_trunk_executor_eval_folds_batchnorm(fake_ops):
    cfg, net, batch = _net_and_batch(seed=2)
    net.eval()
    with torch.no_grad():
        a = _run(net, batch, False)
        b = _run(net, batch, True)
    for h in cfg.network_heads:
        assert torch.allclose(a[h].F, b[h].F, rtol=1e-3, atol=1e-4), (h, float((a[h].F - b[h].F).abs().max()))
    # the folded scale / shift are cached and refreshed when the statistics change
    ex = net.trunk_executor()
    from box2mask_b200.trunk import _fold_bn
    s0, _ = _fold_bn(net.bn0)
    assert _fold_bn(net.bn0)[0] is s0
    with torch.no_grad():
        net.bn0.bn.running_var.mul_(2.0)
    assert _fold_bn(net.bn0)[0] is not s0
    assert len(ex.program.steps) == 81 + 7
