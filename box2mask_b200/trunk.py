"""Trunk executor: the U-Net part of SelectionNet as ONE autograd node with a hand-scheduled forward and backward.

The reference runs the trunk module by module through autograd (/root/reference/models/detection_net.py:235-337,
models/resnet.py:70-83): ~250 modules, ~540 kernel calls per training step, and on the round-1 build 31 ms of Python
dispatch per step around 40 ms of GPU work. The trunk is a STATIC program - 81 units of
convolution -> BatchNorm (+ residual) (+ ReLU) and 7 channel concatenations - so this module compiles it once into a
step list and executes it directly against the C-ABI ops:

  training forward   conv (column statistics in its epilogue) -> one BatchNorm(+residual)(+ReLU) pass per unit
  training backward  the reverse list: BatchNorm backward -> wgrad straight into the flat gradient buffer -> dgrad, where
                     the gradient that is already pending for the unit's input (residual branch, skip connection) is
                     added in the dgrad epilogue instead of by a separate elementwise pass
  eval forward       ONE kernel per unit: eval-mode BatchNorm folded into per-column scale/shift of the convolution
                     epilogue, with the residual add and the ReLU ("BatchNorm and ReLU are fused into the epilogue")

Gradients of the trunk parameters live in one flat fp32 buffer in order of completion during backward (deep levels
first... see TrunkProgram.param_order), which is what lets data-parallel training start the all-reduce of the deep
levels' gradients (>90 % of the bytes) on a side stream while the full-resolution encoder layers are still running
(box2mask_b200/grad_sync.py: TrunkGradSync).

Numerics are those of the module-by-module path (box2mask_b200/me/nn.py), which stays available as the drop-in surface
and is what the reference's own SelectionNet class runs on; tests compare the two.
"""
import torch

from . import ops, peer
from .me.nn import _dgrad_mode, _round16


class Unit:
    """conv -> BatchNorm (+ residual) (+ ReLU). src / dst / res are tensor ids of the program."""
    __slots__ = ("name", "conv", "bn", "src", "dst", "res", "relu", "map", "need_dx", "level", "fuse_for")

    def __init__(self, name, conv, bn, src, dst, res, relu, map_key, need_dx, level):
        self.name, self.conv, self.bn, self.src, self.dst, self.res, self.relu = name, conv, bn, src, dst, res, relu
        self.map, self.need_dx, self.level = map_key, need_dx, level
        self.fuse_for = None      # index of the unit whose BatchNorm-backward reduction this unit's dgrad epilogue takes


class Cat:
    __slots__ = ("a", "b", "dst", "ca", "cb")

    def __init__(self, a, b, dst, ca, cb):
        self.a, self.b, self.dst, self.ca, self.cb = a, b, dst, ca, cb


class TrunkProgram:
    """The step list of a SelectionNet trunk (built once per network)."""

    def __init__(self, net):
        from .selection_net import DECODER, ENCODER
        self.net = net
        self.steps = []
        self._next = 1                      # tensor id 0 = the network input
        nlev = len(ENCODER)

        def new():
            self._next += 1
            return self._next - 1

        def unit(name, conv, bn, src, res, relu, map_key, level, need_dx=True):
            dst = new()
            self.steps.append(Unit(name, conv, bn, src, dst, res, relu, map_key, need_dx, level))
            return dst

        def stage(name, blocks, x, level):
            for i, blk in enumerate(blocks):
                pre = "%s.%d" % (name, i)
                h = unit(pre + ".conv1", blk.conv1, blk.norm1, x, None, True, ("sub", level, 3), level)
                r = x
                if blk.downsample is not None:
                    r = unit(pre + ".downsample", blk.downsample[0], blk.downsample[1], x, None, False, ("id", level), level)
                x = unit(pre + ".conv2", blk.conv2, blk.norm2, h, r, True, ("sub", level, 3), level)
            return x

        stem = unit("conv0p1s1", net.conv0p1s1, net.bn0, 0, None, True, ("sub", 0, net.conv0p1s1.kernel_size), 0, need_dx=False)
        out, skips = stem, []
        for l, (conv, bn, block, _) in enumerate(ENCODER):
            out = unit(conv, getattr(net, conv), getattr(net, bn), out, None, True, ("down", l), l + 1)
            out = stage(block, getattr(net, block), out, l + 1)
            skips.append(out)
        level = nlev
        for conv, bn, block, width, skip in DECODER:
            level -= 1
            up = unit(conv, getattr(net, conv), getattr(net, bn), out, None, True, ("up", level), level)
            other = stem if skip < 0 else skips[skip]
            cb = net.conv0p1s1.out_channels if skip < 0 else ENCODER[skip][3]
            dst = new()
            self.steps.append(Cat(up, other, dst, width, cb))
            out = stage(block, getattr(net, block), dst, level)
        self.out_id = out
        self.units = [s for s in self.steps if isinstance(s, Unit)]
        # Parameters in order of gradient completion during backward = reverse program order. `deep_units`: the units
        # whose gradients are final once backward has left tensor-stride level DEEP_LEVEL on the encoder side.
        self.param_order = []
        for u in reversed(self.units):
            self.param_order += [u.conv.kernel, u.bn.bn.weight, u.bn.bn.bias]
        self._plan_fused_reductions()

    def _plan_fused_reductions(self):
        """BatchNorm backward needs (sum g, sum g * xhat) over the complete gradient g of the layer's output. Where the
        LAST contribution to that gradient in backward order is a dgrad (its epilogue adds whatever is already pending:
        residual branch, skip connection), that dgrad's epilogue takes the reduction too and the layer's own reduction
        pass - a full read of g and x - is dropped. Not fusable: tensors whose last contribution is a concatenation
        split (the transposed convolutions of the decoder) and the trunk output (its gradient comes from the heads)."""
        producer = {u.dst: i for i, u in enumerate(self.units)}
        index = {id(u): i for i, u in enumerate(self.units)}
        last = {}                                     # tensor id -> (kind, unit index) of the last contribution
        for st in reversed(self.steps):
            if isinstance(st, Cat):
                last[st.a] = last[st.b] = ("cat", None)
                continue
            if st.res is not None:
                last[st.res] = ("res", None)
            if st.need_dx:
                last[st.src] = ("dgrad", index[id(st)])
        for tid, (kind, ui) in last.items():
            if kind == "dgrad" and tid in producer and _round16(self.units[ui].conv.out_channels) >= 32:
                self.units[ui].fuse_for = producer[tid]

    def parameters(self):
        return self.param_order


def _fold_bn(bn):
    """Eval-mode BatchNorm as per-column (scale, shift): y = x * scale + shift. Cached on the module, refreshed when
    any of its tensors changes (version counters)."""
    b = bn.bn
    key = (b.weight._version, b.bias._version, b.running_mean._version, b.running_var._version, b.weight.data_ptr())
    cached = bn.__dict__.get("_b2m_fold")
    if cached is None or cached[0] != key:
        with torch.no_grad():
            scale = (b.weight.float() * torch.rsqrt(b.running_var.float() + b.eps)).contiguous()
            shift = (b.bias.float() - b.running_mean.float() * scale).contiguous()
        cached = (key, scale, shift)
        bn.__dict__["_b2m_fold"] = cached
    return cached[1], cached[2]


class _Maps:
    """Kernel maps of one batch by program key."""

    def __init__(self, cm):
        self.cm = cm

    def get(self, key):
        """-> (forward KernelMap or None, backward KernelMap or None, n_out)"""
        cm, kind = self.cm, key[0]
        if kind == "sub":
            km = cm.submanifold_map(2 ** key[1], key[2])
            return km, km, km.n_out
        if kind == "id":
            return None, None, cm.levels[2 ** key[1]].shape[0]
        down, up = cm.stride2_maps(2 ** key[1])
        if kind == "down":
            return down, up, down.n_out
        return up, down, up.n_out


def _packed(conv):
    pre = conv.__dict__.get("_prepacked")
    if pre is None or pre[2] != conv.kernel._version:
        raise RuntimeError("trunk executor: convolution weights are not packed for this step (prepack_conv_weights)")
    return pre[0], pre[1]


def _sync_group(bn, training):
    g = getattr(bn, "process_group", None)
    if g is None or not training or not torch.distributed.is_initialized():
        return None
    if g == "default":
        g = torch.distributed.group.WORLD
        bn.process_group = g
    return g


class GradBuffer:
    """Flat fp32 gradient buffer of the trunk parameters, laid out in order of completion during backward; each
    parameter's `.grad` is pointed at its slice after backward (no copies, like grad_sync.FlatGradSync)."""

    def __init__(self, program):
        params = program.parameters()
        dev = params[0].device
        sizes = [p.numel() for p in params]
        self.params = params
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(torch.split(self.flat, sizes), params)]
        self.index = {id(p): i for i, p in enumerate(params)}
        self.offsets = [0]
        for n in sizes:
            self.offsets.append(self.offsets[-1] + n)

    def view(self, p):
        return self.views[self.index[id(p)]]

    def end_of(self, param):
        """flat offset just past `param` (everything before it in completion order included)"""
        return self.offsets[self.index[id(param)] + 1]

    def stale(self):
        return any(p.device != self.flat.device for p in self.params)


class TrunkFn(torch.autograd.Function):
    """feats bf16/f32 [N0, C_in] -> trunk output bf16 [N0, 96]. `anchor` is a dummy differentiable scalar that keeps
    this node in the graph (the trunk parameters are reached through `ex`, not through autograd inputs: their
    gradients are written into ex.grads and attached to the parameters by backward)."""

    @staticmethod
    def forward(ctx, anchor, feats, ex, cm):
        ctx.ex, ctx.cm = ex, cm
        out, ctx.saved = ex.forward_train(feats, cm)
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.ex.backward(ctx.saved, ctx.cm, dout.contiguous())
        ctx.saved = None
        return None, None, None, None


class TrunkExecutor:
    def __init__(self, net):
        self.net = net
        self.program = TrunkProgram(net)
        self.grads = None
        self.anchor = None
        # data-parallel hooks (grad_sync.TrunkGradSync): on_bucket() fires inside backward right after the unit named
        # bucket_after (every gradient of the flat buffer up to grads.end_of(bucket_after) is then enqueued), on_done()
        # after the last unit
        self.on_bucket = None
        self.bucket_after = None
        self.on_done = None
        # Weight gradients run on a side stream: they are off the critical path of backward (nothing in the pass reads
        # them), while the dgrad chain through tensor-stride levels >= 2 is a sequence of small launches that leaves
        # most SMs idle. The wgrads of the full-resolution decoder stages (levels <= DEFER_LEVEL, ~6 ms of full-GPU
        # work at the start of backward) are held back until the chain enters the deep levels and then fill those SMs.
        self.overlap_wgrad = True
        self._wgrad_stream = None
        # BatchNorm-backward reductions taken by the dgrad that completes the gradient: True = where that dgrad runs
        # offset-split (deep levels), "all" = everywhere (slower, kept for the tests), False = never
        self.fuse_bn_reduce = True

    # ---------------------------------------------------------------------------------------------
    def _input(self, feats, unit):
        if feats.dtype == torch.float32:
            return ops.cast_pad_bf16(feats.contiguous(), _round16(unit.conv.in_channels))
        if feats.dtype != torch.bfloat16:
            raise TypeError("features must be float32 or bfloat16")
        return feats.contiguous()

    def run(self, x):
        """x: SparseTensor input -> features of the last full-resolution stage, bf16 [N0, 96]."""
        net = self.net
        cm = x.coordinate_manager
        training = net.training and torch.is_grad_enabled()
        if not net.training:
            return self.forward_eval(x.F, cm)
        if not training:
            out, _ = self.forward_train(x.F, cm)     # training-mode BatchNorm without a graph (validation in train mode)
            return out
        if self.grads is None or self.grads.stale():
            self.grads = GradBuffer(self.program)
        if self.anchor is None or self.anchor.device != x.F.device:
            self.anchor = torch.zeros((), device=x.F.device, requires_grad=True)
        return TrunkFn.apply(self.anchor, x.F, self, cm)

    # ---------------------------------------------------------------------------------------------
    def forward_eval(self, feats, cm):
        prog, maps = self.program, _Maps(cm)
        T = {0: self._input(feats, prog.units[0])}
        last_use = self._last_use()
        folded = [_fold_bn(st.bn) for st in prog.units]      # (eager torch ops when stale: before any launch is recorded)
        ll = ops.LaunchList.begin(T[0].device)               # the whole pass is issued by one C call (ops.LaunchList)
        try:
            ui = 0
            for i, st in enumerate(prog.steps):
                if isinstance(st, Cat):
                    T[st.dst] = self._cat(ll, T[st.a], T[st.b])
                else:
                    km_f, _, n_out = maps.get(st.map)
                    scale, shift = folded[ui]
                    ui += 1
                    res = T[st.res] if st.res is not None else None
                    T[st.dst] = ops.conv_forward(T[st.src], km_f, _packed(st.conv)[0], st.conv.kernel_volume, n_out,
                                                 st.conv.out_channels, None, scale, shift, res, st.relu)
                for tid in last_use.get(i, ()):
                    T.pop(tid, None)
        finally:
            if ll is not None:
                ll.end()
        return T[prog.out_id]

    @staticmethod
    def _cat(ll, a, b):
        """channel concatenation: two column copies recorded in the launch list, torch.cat otherwise"""
        if ll is None:
            return torch.cat([a, b], 1)
        out = torch.empty((a.shape[0], a.shape[1] + b.shape[1]), dtype=a.dtype, device=a.device)
        ops.copy_columns(a, 0, a.shape[1], out, 0)
        ops.copy_columns(b, 0, b.shape[1], out, a.shape[1])
        return out

    def _last_use(self):
        lu = self.__dict__.get("_lu")
        if lu is None:
            last = {}
            for i, st in enumerate(self.program.steps):
                for tid in ((st.a, st.b) if isinstance(st, Cat) else (st.src, st.res)):
                    if tid is not None:
                        last[tid] = i
            lu = {}
            for tid, i in last.items():
                lu.setdefault(i, []).append(tid)
            self._lu = lu
        return lu

    # ---------------------------------------------------------------------------------------------
    def forward_train(self, feats, cm):
        prog, maps = self.program, _Maps(cm)
        T = {0: self._input(feats, prog.units[0])}
        saved = []
        torch._foreach_add_([u.bn.bn.num_batches_tracked for u in prog.units], 1)     # one launch for all 81 counters
        ll = ops.LaunchList.begin(T[0].device)
        try:
            self._forward_train_steps(prog, maps, T, saved, ll)
        finally:
            if ll is not None:
                ll.end()
        return T[prog.out_id], saved

    def _forward_train_steps(self, prog, maps, T, saved, ll):
        for st in prog.steps:
            if isinstance(st, Cat):
                T[st.dst] = self._cat(ll, T[st.a], T[st.b])
                continue
            conv, b = st.conv, st.bn.bn
            km_f, _, n_out = maps.get(st.map)
            x = T[st.src]
            c_out = conv.out_channels
            dev = x.device
            group = _sync_group(st.bn, True)
            colsum = ops.ZeroArena.take(2 * c_out + 1, dev) if group is not None else ops.ZeroArena.take(2 * c_out, dev)
            y = ops.conv_forward(x, km_f, _packed(conv)[0], conv.kernel_volume, n_out, c_out, colsum[:2 * c_out])
            n_stat, count = n_out, None
            if group is not None:
                # global (sum, sum of squares, row count): one exchange over NVLink peer memory (a launch-list command
                # like the kernels around it), or a torch.distributed all-reduce where the group does not qualify
                sums = peer.allreduce_sum(colsum, group, tail=float(n_out))
                count, n_stat = sums[-1:], 0
            else:
                sums = colsum
                if n_out <= 1:
                    raise ValueError("Expected more than 1 value per channel when training")
            momentum = b.momentum if b.momentum is not None else 1.0 / float(b.num_batches_tracked)
            res = T[st.res] if st.res is not None else None
            out, mean, invstd, mask = ops.bn_forward(y, sums, b.weight.detach(), b.bias.detach(), b.running_mean,
                                                     b.running_var, momentum, b.eps, True, res, st.relu, n_stat, want_mask=True)
            saved.append((x, y, mask, mean, invstd, count, n_stat))
            T[st.dst] = out

    DEFER_LEVEL = 1

    def _side_stream(self, device):
        if not self.overlap_wgrad or device.type != "cuda" or ops.Profile.enabled:
            return None        # (the instrumented pass of bench.py times every launch alone on one stream)
        if self._wgrad_stream is None or self._wgrad_stream.device != device:
            self._wgrad_stream = torch.cuda.Stream(device=device)
        return self._wgrad_stream

    def _wgrad(self, ll, side, ready, x, dx_bn, km_f, conv, n_out, kview):
        """dW of one unit into its slice of the flat gradient buffer, on `side` (after event `ready`) when given.
        With a launch list `ready` is an event slot of the list, otherwise a torch event."""
        padded = kview.shape[-2] != x.shape[1]
        if padded:
            # the 6-channel input padded to 8: wgrad over the padded width, the real rows are copied out by torch - on
            # the main stream (it is the last unit of the pass, nothing is left to overlap with)
            dw = ops.conv_wgrad(x, dx_bn, km_f, conv.kernel_volume, n_out)
            if ll is not None:
                ll.flush()
            kview.copy_(dw[:, :kview.shape[-2], :].reshape(kview.shape))
            return
        if side is None:
            ops.conv_wgrad(x, dx_bn, km_f, conv.kernel_volume, n_out, out=kview)
        elif ll is not None:
            ll.wait_event(ready, stream=1)
            ll.stream = 1
            ops.conv_wgrad(x, dx_bn, km_f, conv.kernel_volume, n_out, out=kview)     # keeps x / dx_bn alive until the flush
            ll.stream = 0
        else:
            side.wait_event(ready)
            with torch.cuda.stream(side):
                ops.conv_wgrad(x, dx_bn, km_f, conv.kernel_volume, n_out, out=kview)
            x.record_stream(side)
            dx_bn.record_stream(side)

    # ---------------------------------------------------------------------------------------------
    def backward(self, saved, cm, dout):
        prog, grads = self.program, self.grads
        side = self._side_stream(dout.device)
        main = torch.cuda.current_stream(dout.device) if side is not None else None
        # gradients are WRITTEN into the flat buffer; anything already accumulated on the parameters (a caller that does
        # not zero the gradients between backward passes) is carried over and added back at the end
        carry = None
        if any(p.grad is not None for p in grads.params):
            carry = torch.zeros_like(grads.flat)
            for p, v in zip(grads.params, torch.split(carry, [q.numel() for q in grads.params])):
                if p.grad is not None:
                    v.view_as(p).copy_(p.grad)
        ll = ops.LaunchList.begin(dout.device, side)      # the pass is issued by one C call per bucket (ops.LaunchList)
        try:
            self._backward_steps(prog, _Maps(cm), grads, saved, dout, ll, side, main)
        finally:
            if ll is not None:
                ll.end()
        if carry is not None:
            grads.flat.add_(carry)
        for p, v in zip(grads.params, grads.views):
            p.grad = v
        if self.on_done is not None:
            self.on_done()

    def _backward_steps(self, prog, maps, grads, saved, dout, ll, side, main):
        deferred, deep_seen = [], False
        G = {prog.out_id: dout}
        R = {}                                  # unit index -> its BatchNorm reduction, taken by a dgrad epilogue
        si = len(saved)

        def join():
            if side is None:
                return
            if ll is not None:
                ll.join_side()
            else:
                main.wait_stream(side)

        for st in reversed(prog.steps):
            if isinstance(st, Cat):
                g = G.pop(st.dst)
                if ll is not None:
                    ga = torch.empty((g.shape[0], st.ca), dtype=g.dtype, device=g.device)
                    gb = torch.empty((g.shape[0], g.shape[1] - st.ca), dtype=g.dtype, device=g.device)
                    ops.copy_columns(g, 0, st.ca, ga, 0)
                    ops.copy_columns(g, st.ca, g.shape[1] - st.ca, gb, 0)
                else:
                    ga, gb = g[:, :st.ca].contiguous(), g[:, st.ca:].contiguous()
                self._acc(G, st.a, ga, ll)
                self._acc(G, st.b, gb, ll)
                continue
            si -= 1
            x, y, mask, mean, invstd, count, n_stat = saved[si]
            saved[si] = None
            conv, b = st.conv, st.bn.bn
            km_f, km_b, n_out = maps.get(st.map)
            g_out = G.pop(st.dst)
            group = _sync_group(st.bn, True)
            hook = None
            if group is not None:
                def hook(red, group=group):
                    return peer.allreduce_sum(red, group)
                hook.deferred = True            # flushes the launch list itself if it has to fall back to a collective
            # the ReLU gate comes from the 1-bit mask the forward pass wrote, not from re-reading the unit's output
            dx_bn, dres, _, _ = ops.bn_backward(y, None, g_out, mean, invstd, b.weight.detach(), st.relu, True,
                                                st.res is not None, n_stat, hook, count,
                                                dgamma=grads.view(b.weight), dbeta=grads.view(b.bias), relu_mask=mask,
                                                red=R.pop(si, None))
            if st.res is not None:
                self._acc(G, st.res, dres, ll)
            kview = grads.view(conv.kernel)
            ready = None
            if side is not None:
                if ll is not None:
                    ready = ll.record_event()
                else:
                    ready = torch.cuda.Event()
                    ready.record(main)
                if st.level > self.DEFER_LEVEL and not deep_seen:
                    deep_seen = True            # the chain enters the deep levels: release the held-back wgrads
                    for job in deferred:
                        self._wgrad(ll, side, *job)
                    deferred = []
            if side is not None and not deep_seen:
                deferred.append((ready, x, dx_bn, km_f, conv, n_out, kview))
            elif side is None or not st.need_dx:
                self._wgrad(ll, side, ready, x, dx_bn, km_f, conv, n_out, kview)
            if st.need_dx:
                # the gradient already pending for this unit's input (residual branch / skip connection) is added in
                # the dgrad epilogue instead of by a separate elementwise pass
                pending = G.get(st.src)
                fused = None
                if st.fuse_for is not None and self.fuse_bn_reduce and (
                        self.fuse_bn_reduce == "all" or
                        ops.conv_forward_is_split(x.shape[0], dx_bn.shape[1], conv.kernel_volume, x.shape[1])):
                    # this dgrad completes the gradient of unit `fuse_for`'s output: its epilogue also takes that unit's
                    # BatchNorm-backward reduction (TrunkProgram._plan_fused_reductions). Only on the offset-split path
                    # of the deep levels, where conv_finalize_kernel owns 8 columns of a row per thread and the sums cost
                    # a few FMAs - and save a launch. In conv_fwd_kernel's epilogue (a lane = a row) the column sums need
                    # two more trips through shared memory per 32-column chunk, and the epilogue warps set the pace:
                    # measured on 1.22 M rows, 96 -> 96: 0.51 -> 0.97 ms, against 0.12 ms for the reduction pass alone.
                    _, py, pmask, pmean, pinvstd, _, _ = saved[st.fuse_for]
                    R[st.fuse_for] = ops.ZeroArena.take(2 * py.shape[1], py.device)
                    fused = (py, pmask, pmean, pinvstd, R[st.fuse_for])
                G[st.src] = ops.conv_forward(dx_bn, km_b, _packed(conv)[1], conv.kernel_volume, x.shape[0], x.shape[1],
                                             residual=pending, bn_reduce=fused)
                if side is not None and deep_seen:
                    # enqueued after the dgrad: the critical path gets the SMs first, the wgrad fills in behind it
                    self._wgrad(ll, side, ready, x, dx_bn, km_f, conv, n_out, kview)
            if self.on_bucket is not None and st.name == self.bucket_after:
                join()                          # the bucket's gradients include wgrads still running on the side stream
                if ll is not None:
                    ll.flush()
                self.on_bucket()
        for job in deferred:
            self._wgrad(ll, side, *job)
        join()

    @staticmethod
    def _acc(G, tid, g, ll=None):
        cur = G.get(tid)
        if cur is not None and ll is not None:
            ll.flush()                          # an eager torch add on tensors that pending launches produce
        G[tid] = g if cur is None else cur.add_(g)
