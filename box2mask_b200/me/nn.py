"""nn.Module wrappers with MinkowskiEngine's names, constructor signatures and parameter names.

Parameter naming is part of the drop-in contract (checkpoint keys such as `block2.0.conv1.kernel`,
`bntr4.bn.bias`, `mlp_semantics.6.kernel`, /root/reference/models/training.py:242-249):
convolutions own `kernel` ([K, C_in, C_out], 2-D when K == 1) and `bias` ([1, C_out]); MinkowskiBatchNorm
owns a torch BatchNorm1d as `.bn`.
"""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .. import ops
from .sparse_tensor import SparseTensor


def _as_int(v):
    if isinstance(v, (list, tuple)):
        if len(set(int(a) for a in v)) != 1:
            raise NotImplementedError("anisotropic kernel/stride is not on the Box2Mask path")
        return int(v[0])
    return int(v)


def _round16(c):
    """bf16 feature width the kernels accept: 8 (the cp.async path of the 6-channel input) or a multiple of 16."""
    return 8 if c <= 8 else (c + 15) // 16 * 16


class _ConvBase(nn.Module):
    is_transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension not in (3, None):
            raise NotImplementedError("only 3 spatial dimensions are on the Box2Mask path")
        if kernel_generator is not None or expand_coordinates:
            raise NotImplementedError("kernel_generator / expand_coordinates are not on the Box2Mask path")
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size, self.stride, self.dilation = _as_int(kernel_size), _as_int(stride), _as_int(dilation)
        if self.dilation != 1:
            raise NotImplementedError("dilation != 1 is not on the Box2Mask path")
        self.kernel_volume = self.kernel_size ** 3
        self.dimension = 3
        shape = (self.in_channels, self.out_channels) if self.kernel_volume == 1 else \
            (self.kernel_volume, self.in_channels, self.out_channels)
        self.kernel = nn.Parameter(torch.empty(shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.empty((1, self.out_channels), dtype=torch.float32)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # MinkowskiEngine default: U(-s, s), s = 1/sqrt(fan * kernel_volume), fan = out for transposed convs
        n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def extra_repr(self):
        return "in=%d, out=%d, kernel_size=%d, stride=%d" % (self.in_channels, self.out_channels, self.kernel_size, self.stride)

    # ---------------------------------------------------------------------------------------------
    def _maps(self, x):
        """-> (nbr_fwd, nbr_bwd, dgrad_mode, n_out, out_stride)"""
        cm, ts = x.coordinate_manager, x._stride
        k, s = self.kernel_size, self.stride
        if not self.is_transpose:
            if k == 1 and s == 1:
                return None, None, 2, len(x), ts
            if s == 1 and k in (3, 5):
                nbr = cm.submanifold_map(ts, k)
                return nbr, nbr, 1, len(x), ts
            if s == 2 and k == 2:
                nbr_down, nbr_up = cm.stride2_maps(ts)
                return nbr_down, nbr_up, 2, nbr_down.n_out, 2 * ts
        else:
            if s == 2 and k == 2:
                if ts % 2 != 0 or (ts // 2) not in cm.stride2:
                    raise RuntimeError("transposed convolution needs the cached finer coordinate map "
                                       "(the encoder's stride-%d level)" % (ts // 2))
                nbr_down, nbr_up = cm.stride2[ts // 2]
                return nbr_up, nbr_down, 2, nbr_up.n_out, ts // 2
        raise NotImplementedError("convolution kernel_size=%d stride=%d transpose=%s is not on the Box2Mask path"
                                  % (k, s, self.is_transpose))

    def forward(self, x):
        if not isinstance(x, SparseTensor):
            raise TypeError("expected a SparseTensor")
        feats = x.F
        if feats.shape[1] != self.in_channels:
            raise RuntimeError("channel mismatch: got %d, expected %d" % (feats.shape[1], self.in_channels))
        if self.kernel_volume == 1 and self.stride == 1 and (feats.dtype == torch.float32 or self.out_channels % 16 != 0):
            # 1x1 convolution on an fp32 tensor (the MLP heads on S superpoint rows, detection_net.py:170-194)
            # or onto a class-count width (final head layer): a plain dense GEMM, done in fp32 by the library.
            out = feats.float() @ self.kernel
            if self.bias is not None:
                out = out + self.bias
            return x._like(out)
        nbr_fwd, nbr_bwd, mode, n_out, out_stride = self._maps(x)
        if feats.dtype == torch.float32:
            feats = ops.cast_pad_bf16(feats.contiguous(), _round16(self.in_channels))
        elif feats.dtype != torch.bfloat16:
            raise TypeError("features must be float32 or bfloat16")
        if feats.shape[1] % 16 != 0 and feats.shape[1] != 8:
            raise NotImplementedError("bf16 feature width must be 8 or a multiple of 16")
        # images packed for this step by prepack_conv_weights (one launch for the whole network), if still current
        pre = getattr(self, "_prepacked", None)
        if pre is not None and (pre[2] != self.kernel._version or pre[3] != feats.shape[1] or pre[4] != mode):
            pre = None
        y, colsum = Fn.SparseConvFn.apply(feats.contiguous(), self.kernel, nbr_fwd, nbr_bwd, mode, n_out, self.in_channels,
                                          pre)
        out = x._like(y, out_stride)
        if self.bias is not None:
            out._F = y + self.bias.to(y.dtype)
        else:
            out._colsum = colsum
            out._colsum_gen = ops.ZeroArena.generation(y.device)
        return out


class MinkowskiConvolution(_ConvBase):
    is_transpose = False


class MinkowskiConvolutionTranspose(_ConvBase):
    is_transpose = True


class MinkowskiBatchNorm(nn.Module):
    """BatchNorm1d over the rows of F; `.bn` holds the parameters and running statistics."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)
        self.process_group = None   # set by MinkowskiSyncBatchNorm

    def forward(self, x, residual=None, relu=False):
        return _bn_act(self, x, residual, relu)

    def __repr__(self):
        b = self.bn
        return "%s(%d, eps=%g, momentum=%g)" % (self.__class__.__name__, b.num_features, b.eps, b.momentum)


def _bn_act(mod, x, residual=None, relu=False):
    bn = mod.bn
    feats = x.F
    if feats.dtype != torch.bfloat16:
        # fp32 rows (the MLP heads on S superpoint rows): torch's BatchNorm1d, fp32
        group = getattr(mod, "process_group", None)
        if group is not None and group != "default" and bn.training and torch.distributed.is_initialized():
            bn.num_batches_tracked += 1
            out = Fn.SyncBatchNormFp32Fn.apply(feats, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                               bn.momentum, bn.eps, group)
        else:
            out = bn(feats)
        if residual is not None:
            out = out + residual.F
        if relu:
            out = torch.relu(out)
        return x._like(out)
    if not (bn.affine and bn.track_running_stats):
        raise NotImplementedError("BatchNorm without affine/running stats is not on the Box2Mask path")
    training = bn.training
    if training:
        bn.num_batches_tracked += 1
    momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
    group = None
    if getattr(mod, "process_group", None) is not None and training and torch.distributed.is_initialized():
        group = mod.process_group
    res = residual.F.contiguous() if residual is not None else None
    colsum = x._colsum if training else None
    if colsum is not None and getattr(x, "_colsum_gen", -1) != ops.ZeroArena.generation(feats.device):
        colsum = None      # the scratch arena was recycled since the convolution ran: recompute the statistics
    out = Fn.BatchNormFn.apply(feats.contiguous(), colsum, bn.weight, bn.bias, bn.running_mean,
                               bn.running_var, momentum, bn.eps, training, res, relu, group)
    return x._like(out)


class MinkowskiSyncBatchNorm(MinkowskiBatchNorm):
    """Batch statistics all-reduced over the process group (reference: models/model.py:25)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, process_group=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)
        self.process_group = process_group if process_group is not None else "default"

    def forward(self, x, residual=None, relu=False):
        if self.process_group == "default":
            self.process_group = torch.distributed.group.WORLD if torch.distributed.is_initialized() else None
        return _bn_act(self, x, residual, relu)

    @classmethod
    def convert_sync_batchnorm(cls, module, process_group=None):
        out = module
        if isinstance(module, MinkowskiBatchNorm) and not isinstance(module, MinkowskiSyncBatchNorm):
            b = module.bn
            out = cls(b.num_features, b.eps, b.momentum, b.affine, b.track_running_stats, process_group)
            out.bn = b
        for name, child in module.named_children():
            if name == "bn" and isinstance(module, MinkowskiBatchNorm):
                continue
            out.add_module(name, cls.convert_sync_batchnorm(child, process_group))
        return out


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._like(torch.relu(x.F))


class _GlobalPool(nn.Module):
    """Pool all rows that share a batch index (column 0 of the coordinates); output row = batch index."""
    mode = "avg"

    def forward(self, x):
        ids = x.C[:, 0].to(torch.int64).contiguous()
        s = int(ids.max().item()) + 1 if ids.numel() else 0
        feats = x.F
        if feats.dtype == torch.float32:
            feats = feats.to(torch.bfloat16)
        fn = Fn.SegmentMeanFn if self.mode == "avg" else Fn.SegmentMaxFn
        out = fn.apply(feats.contiguous(), ids, s)
        coords = torch.zeros((s, 4), dtype=torch.int32, device=out.device)
        coords[:, 0] = torch.arange(s, dtype=torch.int32, device=out.device)
        from .sparse_tensor import CoordinateManager
        return SparseTensor(out, coordinate_manager=CoordinateManager(coords))   # unique by construction


class MinkowskiGlobalAvgPooling(_GlobalPool):
    mode = "avg"


class MinkowskiGlobalMaxPooling(_GlobalPool):
    mode = "max"


def _dgrad_mode(conv):
    """Weight packing of the dgrad operand: 1 = mirrored offsets on the same coordinates (stride-1 k3/k5),
    2 = plain transpose (1x1, strided and transposed convolutions) — what _ConvBase._maps returns."""
    return 1 if (not conv.is_transpose and conv.stride == 1 and conv.kernel_size in (3, 5)) else 2


def prepack_conv_weights(module):
    """Pack the forward and dgrad weight images of every bf16-path convolution under `module` in one launch and
    attach them to the modules (used by their next forward/backward as long as the kernel is not modified).
    Call once per step before the forward pass; convolutions called without it pack their own weights."""
    state = module.__dict__.get("_b2m_packer")
    convs = state[0] if state is not None else None
    if state is None or state[1].stale():
        convs = [m for m in module.modules() if isinstance(m, _ConvBase) and m.kernel.is_cuda and not (
            m.kernel_volume == 1 and m.stride == 1 and (m.bias is not None or m.out_channels % 16 != 0))]
        if not convs:
            return
        jobs = []
        for m in convs:
            c_in = _round16(m.in_channels)
            jobs.append((m.kernel, c_in, 0))         # the live Parameter: WeightPacker.stale() sees re-allocations
            jobs.append((m.kernel, c_in, _dgrad_mode(m)))
        state = (convs, ops.WeightPacker(jobs, convs[0].kernel.device))
        module.__dict__["_b2m_packer"] = state
    packer = state[1]
    versions = tuple(m.kernel._version for m in convs)
    # Inference only: skip the packing when no kernel was modified since the last one. In training mode the images are
    # ALWAYS rebuilt: fused optimizers (torch.optim.Adam(fused=True)) update parameters without bumping `_version`.
    if not module.training and module.__dict__.get("_b2m_packed_versions") == (id(packer), versions):
        return
    packer.run()
    # recorded for eval-mode packings only, so that the first inference pass after training always repacks
    module.__dict__["_b2m_packed_versions"] = None if module.training else (id(packer), versions)
    for i, m in enumerate(convs):
        m.__dict__["_prepacked"] = (packer.buffers[2 * i], packer.buffers[2 * i + 1], m.kernel._version,
                                    _round16(m.in_channels), _dgrad_mode(m))


def conv_bn_act(conv, norm, x, residual=None, relu=True):
    """Fused call used by box2mask_b200's own network: conv (stats in the epilogue) -> BN(+res)(+ReLU)."""
    return _bn_act(norm, conv(x), residual, relu)


def _unsupported_module(name):
    class _Unsupported(nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError("%s is not on the Box2Mask hot path (SURVEY.md §8b)" % name)
    _Unsupported.__name__ = name
    return _Unsupported
