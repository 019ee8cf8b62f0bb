"""Gradient averaging across the scene-sharded ranks (the one real exchange step of the path).

The reference wraps the network in DistributedDataParallel (/root/reference/models/model.py:24): gradients are
averaged over ranks by bucketed NCCL all-reduces launched from autograd hooks, overlapping the backward pass.
On this path that overlap costs more than it hides: the convolution kernels are persistent and hold every SM, so
each NCCL kernel that runs beside them forces a second wave of their CTAs (measured on 2 B200: 45 -> 57-60 ms per
step), while the whole 292 MB fp32 payload is well under a millisecond over NVLink 5 once backward has finished.

FlatGradSync therefore does ONE all-reduce per step, after backward:
  * every parameter gradient is copied into one flat fp32 buffer (a multi-tensor copy, ~0.1 ms for 73 M values),
  * `all_reduce(AVG)` over the flat buffer (NCCL; SUM followed by a division on backends without AVG, e.g. gloo),
  * each `p.grad` is re-pointed at its slice of the buffer (no copy back); the optimizer then reads the slices.
It fires from the autograd engine's end-of-backward callback, queued by the first gradient hook of a backward
pass, so the training loop stays `loss.backward(); optimizer.step()` exactly as in models/training.py:63-70.
Parameters without a gradient in a step contribute zeros (every rank reduces the same layout).
"""
import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        dev = self.params[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in self.params)
        sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(torch.split(self.flat, sizes), self.params)]
        self._queued = False
        self.enabled = True
        self.syncs = 0            # number of all-reduces issued (tests / launch accounting)
        self._avg = dist.get_backend(process_group) == "nccl"
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        # all ranks start from rank 0's parameters and buffers, like DistributedDataParallel's constructor
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group else 0, group=process_group)

    def _on_grad(self, _param):
        if self.enabled and not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)

    @torch.no_grad()
    def _finish(self):
        self._queued = False
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        missing = [v for v, p in zip(self.views, self.params) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if self._avg:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(self.world)
        self.syncs += 1
        for v, p in zip(self.views, self.params):
            p.grad = v

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []



class TrunkGradSync:
    """Gradient averaging for a network whose trunk runs on the trunk executor (box2mask_b200/trunk.py), overlapped
    with the backward pass (the DistributedDataParallel contract of /root/reference/models/model.py:24).

    The executor writes the trunk gradients into one flat fp32 buffer in order of completion. > 90 % of the bytes belong
    to the 256-wide deep levels, and they are complete as soon as backward has climbed back to tensor-stride level 3 on
    the encoder side - while the full-resolution encoder layers (a few milliseconds of convolutions) are still to
    come. So:
      bucket 1  flat[0 : end of `conv4p8s2`]   all-reduced on a side stream from inside backward (on_bucket). While
                it is in flight the persistent convolution kernels are capped to leave `free_sms` SMs to the NCCL
                kernels (b2m_set_option(B2M_OPT_MAX_CTAS)); without the cap each NCCL CTA forces a second wave of
                convolution CTAs (measured in round 1: 45 -> 57 ms per step under DDP).
      bucket 2  the rest of the trunk (full-resolution encoder, a few MB) + the MLP heads, all-reduced at the end of
                backward; the main stream then waits for bucket 1.
    Parameters outside the trunk (the heads) are gathered into a small flat buffer like FlatGradSync does."""

    def __init__(self, net, process_group=None, free_sms=8):
        ex = net.trunk_executor()
        self.ex, self.group = ex, process_group
        self.world = dist.get_world_size(process_group)
        trunk_ids = {id(p) for p in ex.program.parameters()}
        self.other = [p for p in net.parameters() if p.requires_grad and id(p) not in trunk_ids]
        dev = ex.program.parameters()[0].device
        self.dev = dev
        sizes = [p.numel() for p in self.other]
        self.other_flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.other_views = [v.view_as(p) for v, p in zip(torch.split(self.other_flat, sizes), self.other)]
        self._avg = dist.get_backend(process_group) == "nccl"
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.free_sms = int(free_sms)
        self.cut = None
        self.syncs = 0
        self._bucket1_done = None
        self._queued = False
        self.enabled = True
        ex.bucket_after = "conv4p8s2"
        ex.on_bucket = self._bucket1
        ex.on_done = self._trunk_done
        self._handles = [p.register_post_accumulate_grad_hook(self._on_other_grad) for p in self.other]
        with torch.no_grad():     # all ranks start from rank 0's parameters and buffers (DDP's constructor does the same)
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group else 0, group=process_group)

    def _reduce(self, t):
        if self._avg:
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t.div_(self.world)
        self.syncs += 1

    def _queue_finish(self):
        if self.enabled and not self._queued:
            self._queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)

    def _on_other_grad(self, _p):
        self._queue_finish()

    def _bucket1(self):
        """Inside backward, right after the last deep-level unit: everything before `cut` is final."""
        if not self.enabled:
            return
        grads = self.ex.grads
        unit = next(u for u in self.ex.program.units if u.name == self.ex.bucket_after)
        self.cut = grads.end_of(unit.bn.bn.bias)
        if self.comm_stream is None:          # CPU (gloo) tests: no streams, reduce in place right away
            self._reduce(grads.flat[:self.cut])
            return
        from . import _lib
        cur = torch.cuda.current_stream(self.dev)
        self.comm_stream.wait_stream(cur)
        with torch.cuda.stream(self.comm_stream):
            self._reduce(grads.flat[:self.cut])
            self._bucket1_done = torch.cuda.Event()
            self._bucket1_done.record(self.comm_stream)
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        _lib.set_option(_lib.OPT_MAX_CTAS, max(sms - self.free_sms, 1))     # applies to the launches enqueued from here on

    def _trunk_done(self):
        self._queue_finish()

    @torch.no_grad()
    def _finish(self):
        """End of the whole backward pass (autograd engine callback)."""
        self._queued = False
        grads = self.ex.grads
        if self.comm_stream is not None:
            from . import _lib
            _lib.set_option(_lib.OPT_MAX_CTAS, 0)
        if grads is not None and self.cut is not None:
            self._reduce(grads.flat[self.cut:])
        have = [(v, p.grad) for v, p in zip(self.other_views, self.other) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        missing = [v for v, p in zip(self.other_views, self.other) if p.grad is None]
        if missing:
            torch._foreach_zero_(missing)
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if self.other:
            self._reduce(self.other_flat)
            for v, p in zip(self.other_views, self.other):
                p.grad = v
        if self._bucket1_done is not None:
            torch.cuda.current_stream(self.dev).wait_event(self._bucket1_done)
            self._bucket1_done = None
        self.cut = None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
        self.ex.on_bucket = self.ex.on_done = None
