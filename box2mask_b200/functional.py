"""Autograd functions over the C-ABI ops (box2mask_b200/ops.py).

Activations are stored in bf16 (fp32 accumulation inside the kernels); parameters and their gradients
are fp32. Each Function's backward calls the matching CUDA kernels — there is no torch fallback.
"""
import torch

from . import ops


class SparseConvFn(torch.autograd.Function):
    """y = sum_k x[nbr_fwd[k]] @ W[k]  (reference: MinkowskiConvolution / MinkowskiConvolutionTranspose forward,
    /root/reference/models/detection_net.py:235-337; backward = autograd of the same, models/training.py:68).

    nbr_fwd : sorted KernelMap (ops.KernelMap) of the forward relation over n_out rows (None = identity, K == 1)
    nbr_bwd : sorted KernelMap of the transposed relation over n_in rows, used for dgrad
    dgrad_mode : weight packing for dgrad (1 = mirrored offsets, same coordinates; 2 = strided / transposed)
    Returns (y bf16 [n_out, c_out], colsum f64 [2*c_out] = per-column (sum, sum of squares) of y).
    """

    @staticmethod
    def forward(ctx, x, kernel, nbr_fwd, nbr_bwd, dgrad_mode, n_out, c_in_real, prepacked=None):
        """prepacked: optional (forward image, dgrad image) of `kernel` from an ops.WeightPacker run of this step."""
        kvol = 1 if kernel.dim() == 2 else kernel.shape[0]
        c_out = kernel.shape[-1]
        if prepacked is not None:
            packed = prepacked[0]
        else:
            w = kernel.detach()
            if w.shape[-2] != x.shape[1]:  # first conv: input channels zero-padded to the bf16 input width
                pad = x.shape[1] - w.shape[-2]
                w = torch.nn.functional.pad(w, (0, 0, 0, pad))
            packed = ops.pack_weights(w.contiguous(), 0)
        colsum = ops.ZeroArena.take(2 * c_out, x.device)
        y = ops.conv_forward(x, nbr_fwd, packed, kvol, n_out, c_out, colsum)
        ctx.save_for_backward(x, kernel)
        ctx.nbr_fwd, ctx.nbr_bwd, ctx.dgrad_mode, ctx.n_out, ctx.kvol = nbr_fwd, nbr_bwd, dgrad_mode, n_out, kvol
        ctx.c_in_real = c_in_real
        ctx.packed_t = prepacked[1] if prepacked is not None else None
        ctx.mark_non_differentiable(colsum)
        return y, colsum

    @staticmethod
    def backward(ctx, dy, _dcolsum):
        x, kernel = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            packed_t = ctx.packed_t
            if packed_t is None:
                packed_t = ops.pack_weights(kernel.detach().contiguous(), ctx.dgrad_mode)
            dx = ops.conv_forward(dy, ctx.nbr_bwd, packed_t, ctx.kvol, x.shape[0], x.shape[1])
        if ctx.needs_input_grad[1]:
            dw = ops.conv_wgrad(x, dy, ctx.nbr_fwd, ctx.kvol, ctx.n_out)
            if dw.shape[1] != ctx.c_in_real:
                dw = dw[:, :ctx.c_in_real, :]
            dw = dw.reshape(kernel.shape).contiguous()
        return dx, dw, None, None, None, None, None, None


class BatchNormFn(torch.autograd.Function):
    """out = act(BN(x) (+ residual)) over rows; reference: MinkowskiBatchNorm -> BatchNorm1d
    (/root/reference/models/resnet.py:63,66,159), ReLU and `out += residual` (models/resnet.py:67,80-81)."""

    @staticmethod
    def forward(ctx, x, sums, gamma, beta, running_mean, running_var, momentum, eps, training, residual, relu,
                sync_group):
        n = x.shape[0]
        n_stat = n
        count = None
        if training:
            if sums is None:
                sums = ops.colstats(x)
            if sync_group is not None:
                # global (sum, sum of squares, row count) in ONE all-reduce; the count stays on the device
                # (n_stat = 0 tells the kernels to read it from sums[2c]): no host sync per layer
                sums = sync_bn_stats(sums, n, sync_group)
                count = sums[-1:]
                n_stat = 0
            elif n_stat <= 1:
                raise ValueError("Expected more than 1 value per channel when training")
        out, save_mean, save_invstd = ops.bn_forward(x, sums, gamma.detach(), beta.detach(), running_mean, running_var,
                                                     momentum, eps, training, residual, relu, n_stat)
        if count is not None:
            ctx.save_for_backward(x, out, save_mean, save_invstd, gamma, count)
        else:
            ctx.save_for_backward(x, out, save_mean, save_invstd, gamma)
        ctx.relu, ctx.training, ctx.has_res, ctx.n_stat, ctx.sync_group = relu, training, residual is not None, n_stat, sync_group
        return out

    @staticmethod
    def backward(ctx, dout):
        x, out, save_mean, save_invstd, gamma = ctx.saved_tensors[:5]
        count = ctx.saved_tensors[5] if len(ctx.saved_tensors) > 5 else None
        hook = None
        if ctx.sync_group is not None and ctx.training:
            def hook(red):
                # only dx needs the global reduction; bn.weight / bn.bias gradients stay local like torch's
                # SyncBatchNorm (the gradient all-reduce then averages them with every other parameter)
                from .peer import allreduce_sum
                return allreduce_sum(red, ctx.sync_group)
        dx, dres, dgamma, dbeta = ops.bn_backward(x, out, dout.contiguous(), save_mean, save_invstd, gamma.detach(),
                                                  ctx.relu, ctx.training, ctx.has_res, ctx.n_stat, hook, count)
        return dx, None, dgamma, dbeta, None, None, None, None, None, dres, None, None


def sync_bn_stats(sums, n, group):
    """All-reduce the per-column (sum, sum of squares) and the row count over the ranks of `group`: ONE collective
    of 2C+1 doubles (SyncBN forward, models/model.py:25). Returns the packed f64[2C+1] tensor (the count is its last
    element and stays on the device)."""
    packed = torch.empty(sums.numel() + 1, dtype=sums.dtype, device=sums.device)
    packed[:-1] = sums
    if sums.is_cuda and sums.dtype == torch.float64:
        from .peer import allreduce_sum
        return allreduce_sum(packed, group, tail=float(n))      # NVLink peer exchange when the group sits on one node
    packed[-1] = float(n)
    torch.distributed.all_reduce(packed, group=group)
    return packed


class SyncBatchNormFp32Fn(torch.autograd.Function):
    """Synchronised BatchNorm for the small fp32 tensors of the MLP heads (S superpoint rows): plain torch
    arithmetic around two tiny all-reduces; device-agnostic (gloo on CPU in the tests, NCCL on GPUs)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, group):
        xd = x.double()
        sums = sync_bn_stats(torch.cat([xd.sum(0), (xd * xd).sum(0)]), x.shape[0], group)
        n = sums[-1]          # global row count, a 0-d tensor on the device (no host sync)
        c = x.shape[1]
        mean = sums[:c] / n
        var = (sums[c:2 * c] / n - mean * mean).clamp_(min=0)
        invstd = torch.rsqrt(var + eps)
        with torch.no_grad():
            running_mean.mul_(1 - momentum).add_(momentum * mean.to(running_mean.dtype))
            running_var.mul_(1 - momentum).add_(momentum * (var * n / (n - 1).clamp(min=1)).to(running_var.dtype))
        xhat = ((xd - mean) * invstd).to(x.dtype)
        ctx.save_for_backward(xhat, gamma, invstd.to(x.dtype))
        ctx.n, ctx.group = n, group
        return xhat * gamma + beta

    @staticmethod
    def backward(ctx, dout):
        xhat, gamma, invstd = ctx.saved_tensors
        red = torch.cat([dout.sum(0), (dout * xhat).sum(0)]).double()
        dgamma, dbeta = red[xhat.shape[1]:].to(dout.dtype), red[:xhat.shape[1]].to(dout.dtype)
        if red.is_cuda:
            from .peer import allreduce_sum
            red = allreduce_sum(red.contiguous(), ctx.group)
        else:
            torch.distributed.all_reduce(red, group=ctx.group)
        c = xhat.shape[1]
        sg, sgx = (red[:c] / ctx.n).to(dout.dtype), (red[c:] / ctx.n).to(dout.dtype)
        dx = gamma * invstd * (dout - sg - xhat * sgx)
        return dx, dgamma, dbeta, None, None, None, None, None


class SegmentMeanFn(torch.autograd.Function):
    """Superpoint mean pooling (/root/reference/models/detection_net.py:345-352)."""

    @staticmethod
    def forward(ctx, f, ids, s):
        out, counts = ops.segment_mean_forward(f, ids, s)
        ctx.save_for_backward(ids, counts)
        ctx.n = f.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, counts = ctx.saved_tensors
        return ops.segment_mean_backward(dout.contiguous().float(), ids, counts, ctx.n), None, None


class SegmentMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, ids, s):
        out, argmax = ops.segment_max_forward(f, ids, s)
        ctx.save_for_backward(argmax)
        ctx.shape = f.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        (argmax,) = ctx.saved_tensors
        return ops.segment_max_backward(dout.contiguous().float(), argmax, ctx.shape[0]), None, None
