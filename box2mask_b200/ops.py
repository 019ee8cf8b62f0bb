"""Thin functional wrappers over the C-ABI (include/b2m.h). torch is used only for device memory and
streams; every computation below happens in libb2m.so. All tensors must live on a CUDA device."""
import ctypes
import struct

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _lib_or_raise():
    return _lib.load()


class Profile:
    """Launch accounting (always on) and optional CUDA-event timing of every C-ABI call (bench.py)."""
    launches = 0          # kernels launched through the C-ABI since the last reset
    enabled = False       # when True, each call is bracketed by CUDA events on the current stream
    records = []          # (kind, algorithmic flops, algorithmic bytes, start event, end event)
    _pairs = {}           # id(nbr) -> number of valid (in,out) pairs

    @classmethod
    def reset(cls):
        cls.launches, cls.records, cls._pairs = 0, [], {}

    @classmethod
    def pairs(cls, nbr, n_out):
        if nbr is None:
            return n_out
        key = (nbr.data_ptr(), tuple(nbr.shape))
        if key not in cls._pairs:
            cls._pairs[key] = int((nbr >= 0).sum().item())
        return cls._pairs[key]


def _run(kind, n_kernels, call, flops=None, nbytes=None, tag=""):
    Profile.launches += n_kernels
    if not Profile.enabled:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f = flops() if callable(flops) else flops
    b = nbytes() if callable(nbytes) else nbytes
    if callable(tag):
        tag = tag()
    e0.record()
    r = call()
    e1.record()
    Profile.records.append((kind, f or 0, b or 0, e0, e1, tag))
    return r


def _p(t):
    return 0 if t is None else t.data_ptr()


class LaunchList:
    """Recorder for b2m_run_commands (include/b2m.h, "Launch lists"): while one is current, conv_forward / conv_wgrad /
    bn_forward / bn_backward / copy_columns append their C-ABI arguments to a command buffer instead of calling the
    library, and flush() issues the whole pass in one call - ~5 us of launch cost per kernel instead of ~50 us of
    Python, ctypes and stream bookkeeping per call (a training step has ~540 of them).

    Rules that keep deferred launches correct with torch's caching allocator:
      * commands run in recording order on the main stream, so a block that Python frees and torch.empty hands out
        again during recording is reused in stream order, exactly as with immediate launches - as long as no *eager*
        torch kernel runs while commands are pending: call flush() before any torch op on tensors the list touches;
      * tensors read by side-stream commands are kept alive until flush() and then marked with record_stream(side)
        (the allocator must not reuse them for main-stream work before the side stream is done)."""
    current = None
    _S = struct.Struct("<ii28q2d")
    _PAD = (0,) * 28
    OP_CONV_FORWARD, OP_CONV_WGRAD, OP_BN_FORWARD, OP_BN_BACKWARD_REDUCE, OP_BN_BACKWARD_APPLY = 1, 2, 3, 4, 5
    OP_COPY_COLUMNS, OP_RECORD, OP_WAIT, OP_PEER_ALLREDUCE = 6, 7, 8, 9

    FLUSH_EVERY = 8      # commands per b2m_run_commands call: the GPU starts on a pass while the host still records it

    def __init__(self, device, side_stream=None, capacity=64):
        self.device = torch.device(device)
        self.side = side_stream
        self.cap = capacity
        self.buf = bytearray(self._S.size * capacity)
        self.cbuf = (ctypes.c_char * len(self.buf)).from_buffer(self.buf)
        self.n = 0
        self.kinds = []
        self.stream = 0          # stream selector of the commands being recorded (1 = side)
        self.keep_side = []      # tensors read or written by side-stream commands
        self.next_event = 0
        self.flushes = 0

    @classmethod
    def begin(cls, device, side_stream=None):
        """A new current list, or None where launches stay immediate (CPU stand-ins, the instrumented pass)."""
        device = torch.device(device)
        if device.type != "cuda" or Profile.enabled or cls.current is not None:
            return None
        cls.current = cls(device, side_stream)
        return cls.current

    def end(self):
        try:
            self.flush()
        finally:
            if LaunchList.current is self:
                LaunchList.current = None

    def add(self, kind, op, args, f0=0.0, f1=0.0):
        if self.n >= self.FLUSH_EVERY:
            self.flush()
        self._S.pack_into(self.buf, self.n * self._S.size, op, self.stream, *args, *self._PAD[len(args):], f0, f1)
        self.kinds.append(kind)
        self.n += 1

    def target_stream(self):
        """the torch stream the commands being recorded will run on (None = the current stream)"""
        return self.side if self.stream else None

    def record_event(self, stream=0):
        slot = self.next_event
        self.next_event += 1
        cur, self.stream = self.stream, stream
        self.add("record", self.OP_RECORD, (slot,))
        self.stream = cur
        return slot

    def wait_event(self, slot, stream):
        cur, self.stream = self.stream, stream
        self.add("wait", self.OP_WAIT, (slot,))
        self.stream = cur

    def join_side(self):
        """the main stream waits for everything recorded on the side stream so far"""
        self.wait_event(self.record_event(stream=1), stream=0)

    def flush(self):
        n, kinds = self.n, self.kinds
        self.n, self.kinds = 0, []
        try:
            if n:
                self.flushes += 1
                failed = ctypes.c_int64(-1)
                side = self.side.cuda_stream if self.side is not None else None
                code = _lib.load().b2m_run_commands(self.cbuf, n, stream_ptr(self.device), side, ctypes.byref(failed))
                if code != 0:
                    i = failed.value
                    check(code, "launch list command %d of %d (%s)" % (i, n, kinds[i] if 0 <= i < n else "?"))
        finally:
            keep, self.keep_side = self.keep_side, []
            if self.side is not None:
                for t in keep:
                    t.record_stream(self.side)

    @classmethod
    def flush_current(cls):
        if cls.current is not None:
            cls.current.flush()


class ZeroArena:
    """Zeroed fp64 scratch for the per-layer statistics (conv epilogue column sums, BatchNorm backward reductions):
    slices of one pre-zeroed buffer per device instead of one fill kernel per layer (243 per training step).
    When the buffer is used up a fresh zeroed one replaces it (slices still in use keep the old one alive) and
    `generation` advances, which makes holders of cached column sums recompute them."""
    SIZE = 1 << 19      # doubles (4 MB)
    _state = {}         # device -> [buffer, next offset, generation]

    @classmethod
    def take(cls, n, device):
        device = torch.device(device)
        st = cls._state.get(device)
        n_al = (int(n) + 31) // 32 * 32
        if st is None:
            st = [torch.zeros(cls.SIZE, dtype=torch.float64, device=device), 0, 0]
            cls._state[device] = st
        elif st[1] + n_al > cls.SIZE:
            # a fresh zeroed buffer, not zero_() in place: slices handed out earlier may still be waiting for their
            # consumer (a reduction produced by a dgrad epilogue is read several layers later); they keep the old
            # buffer alive. (An eager fill: commands still pending in a launch list go first.)
            LaunchList.flush_current()
            st[0] = torch.zeros(cls.SIZE, dtype=torch.float64, device=device)
            st[1] = 0
            st[2] += 1
        out = st[0][st[1]:st[1] + n]
        st[1] += n_al
        return out

    @classmethod
    def generation(cls, device):
        st = cls._state.get(torch.device(device))
        return st[2] if st is not None else 0


def _cuda(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise _lib.B2MError("%s must be a CUDA tensor (no CPU fallback in the product path)" % name)
    if t.device.index != torch._C._cuda_getDevice():
        # kernels launch on the current device's stream; a tensor of another device would be dereferenced there
        raise _lib.B2MError("%s lives on cuda:%d but the current device is cuda:%d - call torch.cuda.set_device "
                            "(or use `with torch.cuda.device(...)`) first" % (name, t.device.index, torch.cuda.current_device()))
    if dtype is not None and t.dtype != dtype:
        raise _lib.B2MError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.B2MError("%s must be contiguous" % name)
    return t


def as_u16(t):
    """bf16 tensor viewed as the uint16 bit patterns the C-ABI carries (no copy)."""
    return t


# ---------------------------------------------------------------------------------------------
# coordinates
# ---------------------------------------------------------------------------------------------
class HashTable:
    __slots__ = ("keys", "vals", "capacity", "status")

    def __init__(self, keys, vals, capacity, status):
        self.keys, self.vals, self.capacity, self.status = keys, vals, capacity, status


def hash_build(coords):
    lib = _lib_or_raise()
    _cuda(coords, torch.int32, "coords")
    n = coords.shape[0]
    cap = lib.b2m_hash_capacity(n)
    keys = torch.empty(cap, dtype=torch.int64, device=coords.device)
    vals = torch.empty(cap, dtype=torch.int32, device=coords.device)
    status = torch.empty(2, dtype=torch.int32, device=coords.device)
    _run("hash_build", 1, lambda: check(lib.b2m_hash_build(ptr(coords), n, ptr(keys), ptr(vals), cap, ptr(status),
                                                           stream_ptr()), "hash_build"), nbytes=16 * n + 12 * cap)
    return HashTable(keys, vals, cap, status)


def hash_query(table, query):
    lib = _lib_or_raise()
    _cuda(query, torch.int32, "query")
    rows = torch.empty(query.shape[0], dtype=torch.int32, device=query.device)
    check(lib.b2m_hash_query(ptr(query), query.shape[0], ptr(table.keys), ptr(table.vals), table.capacity, ptr(rows),
                             stream_ptr()), "hash_query")
    return rows


def downsample_coords(coords, new_stride):
    """-> (out_coords int32[M,4] sorted unique, parent_row int32[N]). One host sync to read M."""
    lib = _lib_or_raise()
    _cuda(coords, torch.int32, "coords")
    n = coords.shape[0]
    out = torch.empty((n, 4), dtype=torch.int32, device=coords.device)
    parent = torch.empty(n, dtype=torch.int32, device=coords.device)
    n_out = torch.zeros(1, dtype=torch.int32, device=coords.device)
    ws_bytes = lib.b2m_downsample_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=coords.device)
    _run("downsample_coords", 8, lambda: check(lib.b2m_downsample_coords(
        ptr(coords), n, int(new_stride), ptr(out), ptr(parent), ptr(n_out), ptr(ws), ws_bytes, stream_ptr()),
        "downsample_coords"), nbytes=16 * n + 4 * n)
    m = int(n_out.item())
    return out[:m].contiguous(), parent


def map_pitch(n):
    """Row pitch of every neighbour table / `order` array over n rows (b2m_map_pitch): n rounded up to 128."""
    return (int(n) + 127) // 128 * 128


def kernel_map_submanifold(coords, tensor_stride, kernel_size, table):
    """-> nbr int32[K^3, pitch(n)] (columns >= n hold -1)."""
    lib = _lib_or_raise()
    n = coords.shape[0]
    nbr = torch.empty((kernel_size ** 3, map_pitch(n)), dtype=torch.int32, device=coords.device)
    _run("kernel_map_submanifold", 1, lambda: check(lib.b2m_kernel_map_submanifold(
        ptr(coords), n, int(tensor_stride), int(kernel_size), ptr(table.keys), ptr(table.vals), table.capacity, ptr(nbr),
        stream_ptr()), "kernel_map_submanifold"), nbytes=16 * n + 4 * n * kernel_size ** 3)
    return nbr


def kernel_map_from_coarse(coords, tensor_stride, kernel_size, parent_row, nbr3_coarse, nbr_down, n_coarse, want_gmask=False):
    """The table of kernel_map_submanifold built without hashing from the next coarser level's (unsorted) 3^3 table and
    child table (b2m_kernel_map_from_coarse). -> nbr int32[K^3, pitch(n)] (and gmask int32[groups, words] of the unsorted
    table when want_gmask)."""
    lib = _lib_or_raise()
    _cuda(coords, torch.int32, "coords")
    n = coords.shape[0]
    kvol = kernel_size ** 3
    nbr = torch.empty((kvol, map_pitch(n)), dtype=torch.int32, device=coords.device)
    gmask = torch.empty(((n + 63) // 64, (kvol + 31) // 32), dtype=torch.int32, device=coords.device) if want_gmask else None
    ws_bytes = int(lib.b2m_kernel_map_from_coarse_workspace_bytes(int(n_coarse)))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=coords.device)       # transposed child table
    _run("kernel_map_from_coarse", 2, lambda: check(lib.b2m_kernel_map_from_coarse(
        ptr(coords), n, int(tensor_stride), int(kernel_size), ptr(parent_row), ptr(nbr3_coarse), ptr(nbr_down), int(n_coarse),
        ptr(nbr), ptr(gmask), ptr(ws), ws_bytes, stream_ptr()), "kernel_map_from_coarse"), nbytes=20 * n + 4 * n * kvol)
    return (nbr, gmask) if want_gmask else nbr


def kernel_map_stride2(fine_coords, parent_row, n_coarse, fine_stride):
    lib = _lib_or_raise()
    n_fine = fine_coords.shape[0]
    nbr_down = torch.empty((8, map_pitch(n_coarse)), dtype=torch.int32, device=fine_coords.device)
    nbr_up = torch.empty((8, map_pitch(n_fine)), dtype=torch.int32, device=fine_coords.device)
    _run("kernel_map_stride2", 1, lambda: check(lib.b2m_kernel_map_stride2(
        ptr(fine_coords), n_fine, ptr(parent_row), n_coarse, int(fine_stride), ptr(nbr_down), ptr(nbr_up), stream_ptr()),
        "kernel_map_stride2"), nbytes=20 * n_fine + 32 * n_fine + 32 * n_coarse)
    return nbr_down, nbr_up


class KernelMap:
    """Sorted kernel map consumed by the convolutions: nbr int32[K, pitch(n_out)] (already permuted), order
    int32[pitch(n_out)] (position -> output row, -1 in the padding), gmask uint32-as-int32 [groups, words].
    `raw` keeps the unsorted table (tests)."""
    __slots__ = ("nbr", "order", "gmask", "kvol", "n_out", "raw")

    def __init__(self, nbr, order, gmask, kvol, n_out, raw=None):
        self.nbr, self.order, self.gmask, self.kvol, self.n_out, self.raw = nbr, order, gmask, kvol, n_out, raw


SORT_BLOCK_ROWS = 32768
SORT_MIN_ROWS = 8192          # maps with fewer output rows keep their row order (masks are still produced)


def pad_table(nbr):
    """Unpadded table int32[K, n] (e.g. built on the host by a test) -> int32[K, pitch(n)] with -1 padding."""
    n = nbr.shape[1]
    return torch.nn.functional.pad(nbr, (0, map_pitch(n) - n), value=-1).contiguous()


def sort_kernel_map(nbr, n_out=None, block_rows=None, keep_raw=False):
    """nbr int32[K, pitch(n_out)] from kernel_map_submanifold / kernel_map_stride2 -> KernelMap.
    With n_out=None the table is taken as unpadded int32[K, n_out] and padded first."""
    lib = _lib_or_raise()
    _cuda(nbr, torch.int32, "nbr")
    if n_out is None:
        n_out = nbr.shape[1]
        nbr = pad_table(nbr)
    kvol, pitch = nbr.shape
    if pitch != map_pitch(n_out):
        raise _lib.B2MError("neighbour table must have the padded pitch %d, got %d" % (map_pitch(n_out), pitch))
    if block_rows is None and n_out < SORT_MIN_ROWS:
        # a level of a few thousand rows is a few dozen MMA tiles that each contain (almost) every offset whatever the
        # order: the five launches of a sort buy nothing there (12 of the 22 maps of a ScanNet-shape step)
        block_rows = 0
    if block_rows is None:
        # blocks of >= 32768 rows (their features stay L2-resident while a block is swept); doubled while that lets the
        # (block, mask) sort key fit 32 bits (one radix pass fewer, half the key bytes)
        block_rows = SORT_BLOCK_ROWS
        while kvol <= 31 and block_rows < 262144 and (n_out + block_rows - 1) // block_rows > (1 << (32 - kvol)):
            block_rows *= 2
    words = (kvol + 31) // 32
    order = torch.empty(pitch, dtype=torch.int32, device=nbr.device)
    do_sort = block_rows > 0 and kvol <= 32
    nbr_sorted = torch.empty_like(nbr) if do_sort else nbr
    gmask = torch.empty(((n_out + 63) // 64, words), dtype=torch.int32, device=nbr.device)
    ws_bytes = lib.b2m_kernel_map_sort_workspace_bytes(n_out) if do_sort else 0
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=nbr.device)
    _run("kernel_map_sort", 5, lambda: check(lib.b2m_kernel_map_sort(
        ptr(nbr), kvol, n_out, int(block_rows), ptr(order), ptr(nbr_sorted), ptr(gmask), ptr(ws), ws_bytes, stream_ptr()),
        "kernel_map_sort"), nbytes=12 * n_out * kvol + 16 * n_out)
    return KernelMap(nbr_sorted, order, gmask, kvol, n_out, nbr if keep_raw else None)


def kernel_map_count(nbr, n_out):
    """pairs per offset of a padded table int32[K, pitch(n_out)]."""
    lib = _lib_or_raise()
    counts = torch.empty(nbr.shape[0], dtype=torch.int32, device=nbr.device)
    check(lib.b2m_kernel_map_count(ptr(nbr), nbr.shape[0], int(n_out), ptr(counts), stream_ptr()), "kernel_map_count")
    return counts


# ---------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------
def cast_pad_bf16(x, c_pad):
    lib = _lib_or_raise()
    _cuda(x, torch.float32, "x")
    out = torch.empty((x.shape[0], c_pad), dtype=torch.bfloat16, device=x.device)
    _run("cast_pad_bf16", 1, lambda: check(lib.b2m_cast_pad_bf16(ptr(x), x.shape[0], x.shape[1], c_pad, ptr(out),
                                                               stream_ptr()), "cast_pad_bf16"),
         nbytes=4 * x.numel() + 2 * x.shape[0] * c_pad)
    return out


def pack_weights(kernel, mode):
    """kernel f32 [K, C_in, C_out] (or [C_in, C_out]) -> packed bf16 UMMA B image (uint8 buffer)."""
    lib = _lib_or_raise()
    _cuda(kernel, torch.float32, "kernel")
    if kernel.dim() == 2:
        kvol, (c_in, c_out) = 1, kernel.shape
    else:
        kvol, c_in, c_out = kernel.shape
    nbytes = lib.b2m_packed_weight_bytes(kvol, c_in, c_out, mode)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=kernel.device)
    _run("pack_weights", 1, lambda: check(lib.b2m_pack_weights(ptr(kernel), kvol, c_in, c_out, mode, ptr(packed),
                                                             stream_ptr()), "pack_weights"),
         nbytes=kvol * c_in * c_out * 4 + nbytes)
    return packed


class WeightPacker:
    """Packs the kernels of many convolutions in ONE launch (b2m_pack_weights_batched). jobs: list of
    (kernel f32 [K, c_in_src, c_out] or [c_in_src, c_out], c_in (>= c_in_src, the padded bf16 input width), mode);
    buffers[i] is the packed image of job i, refreshed by run()."""

    def __init__(self, jobs, device):
        lib = _lib_or_raise()
        # jobs may carry Parameters (preferred) or plain tensors: the live objects are kept, not detached aliases, so
        # that a re-allocation by module.to() / load_state_dict(assign=True) is seen by stale()
        self.kernels = [j[0] for j in jobs]
        self.device = torch.device(device)
        self.buffers, meta, prefix = [], [], [0]
        for kernel, c_in, mode in jobs:
            _cuda(kernel, torch.float32, "kernel")
            kvol = 1 if kernel.dim() == 2 else kernel.shape[0]
            c_in_src, c_out = kernel.shape[-2], kernel.shape[-1]
            nbytes = lib.b2m_packed_weight_bytes(kvol, c_in, c_out, mode)
            self.buffers.append(torch.empty(nbytes, dtype=torch.uint8, device=device))
            meta.append([kvol, c_in_src, c_in, c_out, mode])
            prefix.append(prefix[-1] + nbytes // 16)
        self.total = prefix[-1]
        self.n = len(jobs)
        self.ptrs = tuple(k.data_ptr() for k in self.kernels)
        self.src = torch.tensor(self.ptrs, dtype=torch.int64, device=device)
        self.dst = torch.tensor([b.data_ptr() for b in self.buffers], dtype=torch.int64, device=device)
        self.meta = torch.tensor(meta, dtype=torch.int32, device=device)
        self.prefix = torch.tensor(prefix, dtype=torch.int64, device=device)

    def stale(self):
        """True when a parameter was re-allocated or moved (e.g. `.to(device)`): the job table must be rebuilt."""
        return any(k.device != self.device for k in self.kernels) or \
            tuple(k.data_ptr() for k in self.kernels) != self.ptrs

    def run(self):
        lib = _lib_or_raise()
        _run("pack_weights", 1, lambda: check(lib.b2m_pack_weights_batched(
            ptr(self.src), ptr(self.dst), ptr(self.meta), ptr(self.prefix), self.n, self.total, stream_ptr()),
            "pack_weights_batched"), nbytes=sum(k.numel() * 4 for k in self.kernels) + 16 * self.total)


class Workspace:
    """Scratch for the offset-split convolutions (fp32 partial tiles of the deep levels): one growing buffer per
    (device, stream); a call's partials are consumed by the finalize kernel launched inside the same C-ABI call, so
    stream order makes reuse by the next call safe."""
    _buf = {}

    @classmethod
    def get(cls, nbytes, device, stream=None):
        """stream: the torch stream the consuming launches run on (default: the current stream). The buffer is allocated
        under that stream: a block the caching allocator recycles from another stream's pool may still be in use by
        launches of that stream."""
        device = torch.device(device)
        key = (device, stream_ptr(device) if stream is None else stream.cuda_stream)
        buf = cls._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and LaunchList.current is not None:
                LaunchList.current.keep_side.append(buf)      # pending commands may still use the old buffer
            size = max(int(nbytes), 1 << 22)
            if stream is None:
                buf = torch.empty(size, dtype=torch.uint8, device=device)
            else:
                with torch.cuda.stream(stream):
                    buf = torch.empty(size, dtype=torch.uint8, device=device)
            cls._buf[key] = buf
        return buf


def conv_forward_is_split(n_out, c_red, kvol, c_n):
    """True when b2m_conv_forward runs this shape offset-split (few row tiles: the kernel offsets are dealt to several
    CTAs per tile and conv_finalize_kernel sums the slices and runs the epilogue)."""
    return _lib_or_raise().b2m_conv_forward_workspace_bytes(n_out, c_red, kvol, c_n) > 0


def conv_forward(x, kmap, packed_w, kvol, n_out, c_n, colsum=None, scale=None, shift=None, residual=None, relu=False,
                 out_fp32_cols=None, bn_reduce=None):
    """y bf16[n_out, c_n] = epilogue(sum_k x[nbr[k]] @ B[k]) over a sorted KernelMap (None = identity, kvol 1).
    Epilogue (optional): * scale[c_n] + shift[c_n] (+ residual bf16[n_out, c_n]) (ReLU). colsum f64[2*c_n] (zeroed by
    the caller) accumulates the statistics of the result. out_fp32_cols = c: return fp32 [n_out, c] (the first c
    columns) instead of bf16.
    bn_reduce = (x_prod bf16[n_out, c_n], relu_mask uint8[n_out, c_n/8] or None, mean f32[c_n], invstd f32[c_n],
    red f64[2*c_n] zeroed): the call is a dgrad and its result the complete gradient of a BatchNorm layer's output; the
    epilogue also accumulates that layer's backward reduction (sum g, sum g*xhat) into red (b2m_conv_dgrad_bn_reduce), which
    bn_backward(red=...) then takes instead of running its reduction pass."""
    lib = _lib_or_raise()
    _cuda(x, torch.bfloat16, "x")
    bx, bmask, bmean, binv, bred = bn_reduce if bn_reduce is not None else (None, None, None, None, None)
    nbr = kmap.nbr if kmap is not None else None
    order = kmap.order if kmap is not None else None
    gmask = kmap.gmask if kmap is not None else None
    y = y32 = None
    if out_fp32_cols is None:
        y = torch.empty((n_out, c_n), dtype=torch.bfloat16, device=x.device)
    else:
        y32 = torch.empty((n_out, int(out_fp32_cols)), dtype=torch.float32, device=x.device)
    ws_bytes = lib.b2m_conv_forward_workspace_bytes(n_out, x.shape[1], kvol, c_n)
    ll = LaunchList.current
    if ll is not None:
        ws = Workspace.get(ws_bytes, x.device, ll.target_stream()) if ws_bytes else None
        Profile.launches += 2 if ws_bytes else 1
        ll.add("conv_forward", ll.OP_CONV_FORWARD, (
            x.data_ptr(), x.shape[0], x.shape[1], _p(nbr), _p(order), _p(gmask), kvol, n_out, packed_w.data_ptr(), c_n,
            _p(y), _p(colsum), _p(scale), _p(shift), _p(residual), 1 if relu else 0, _p(y32),
            int(out_fp32_cols) if out_fp32_cols is not None else 0, _p(ws), ws_bytes,
            _p(bx), _p(bmask), _p(bmean), _p(binv), _p(bred)))
        return y if y32 is None else y32
    ws = Workspace.get(ws_bytes, x.device) if ws_bytes else None
    _run("conv_forward", 2 if ws_bytes else 1, lambda: check(lib.b2m_conv_dgrad_bn_reduce(
        ptr(x), x.shape[0], x.shape[1], ptr(nbr), ptr(order), ptr(gmask), kvol, n_out, ptr(packed_w), c_n, ptr(y),
        ptr(colsum), ptr(scale), ptr(shift), ptr(residual), int(bool(relu)), ptr(y32),
        int(out_fp32_cols) if out_fp32_cols is not None else 0, ptr(ws), ws_bytes, ptr(bx), ptr(bmask), ptr(bmean),
        ptr(binv), ptr(bred), stream_ptr()), "conv_forward"),
        flops=lambda: 2.0 * Profile.pairs(nbr, n_out) * x.shape[1] * c_n,
        nbytes=lambda: 2.0 * Profile.pairs(nbr, n_out) * x.shape[1] + 2.0 * n_out * c_n,
        tag=lambda: "k%d %d->%d n_in=%d n_out=%d" % (kvol, x.shape[1], c_n, x.shape[0], n_out))
    return y if y32 is None else y32


def conv_wgrad(x, dy, kmap, kvol, n_out, out=None):
    """dw f32 [kvol, c_in, c_out] (overwritten). out: optional destination with kvol * c_in * c_out contiguous floats
    (e.g. a slice of a flat gradient buffer)."""
    lib = _lib_or_raise()
    _cuda(x, torch.bfloat16, "x")
    _cuda(dy, torch.bfloat16, "dy")
    nbr = kmap.nbr if kmap is not None else None
    order = kmap.order if kmap is not None else None
    gmask = kmap.gmask if kmap is not None else None
    c_in, c_out = x.shape[1], dy.shape[1]
    if out is not None:
        if out.numel() != kvol * c_in * c_out or out.dtype != torch.float32 or not out.is_contiguous():
            raise _lib.B2MError("conv_wgrad: out must hold kvol * c_in * c_out contiguous floats")
        dw = out
    else:
        dw = torch.empty((kvol, c_in, c_out), dtype=torch.float32, device=x.device)    # overwritten (zero-filled inside if needed)
    # partial-sum workspace for the row splits (deterministic reduction instead of atomics); allocated on the current
    # stream by the caching allocator
    ws_bytes = int(lib.b2m_conv_wgrad_workspace_bytes(n_out, c_in, c_out, kvol))
    ll = LaunchList.current
    if ll is not None:
        # launches of one stream run one after the other: they share that stream's workspace
        ws = Workspace.get(ws_bytes, x.device, ll.target_stream()) if ws_bytes else None
        Profile.launches += 2 if ws is not None else 1
        ll.add("conv_wgrad", ll.OP_CONV_WGRAD, (x.data_ptr(), x.shape[0], c_in, dy.data_ptr(), c_out, _p(nbr), _p(order),
                                                 _p(gmask), kvol, n_out, dw.data_ptr(), _p(ws), ws_bytes))
        if ll.stream:
            ll.keep_side += (x, dy, dw)
        return dw
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
    _run("conv_wgrad", 2 if ws is not None else 1, lambda: check(lib.b2m_conv_wgrad_ex(
        ptr(x), x.shape[0], c_in, ptr(dy), c_out, ptr(nbr), ptr(order), ptr(gmask), kvol, n_out, ptr(dw), ptr(ws), ws_bytes,
        stream_ptr()), "conv_wgrad"),
        flops=lambda: 2.0 * Profile.pairs(nbr, n_out) * c_in * c_out,
        nbytes=lambda: 2.0 * Profile.pairs(nbr, n_out) * (c_in + c_out),
        tag=lambda: "k%d %d->%d n_in=%d n_out=%d" % (kvol, c_in, c_out, x.shape[0], n_out))
    return dw


def copy_columns(src, src_col, width, dst, dst_col):
    """dst[:, dst_col:dst_col+width] = src[:, src_col:src_col+width] for bf16 [n, c] tensors (column offsets and width
    multiples of 8): ME.cat of the decoder and the split of its gradient, without torch's cat / strided-copy kernels."""
    lib = _lib_or_raise()
    n = src.shape[0]
    if dst.shape[0] != n or src.dtype != torch.bfloat16 or dst.dtype != torch.bfloat16:
        raise _lib.B2MError("copy_columns: bf16 tensors with the same number of rows expected")
    if src_col + width > src.shape[1] or dst_col + width > dst.shape[1] or not (src.is_contiguous() and dst.is_contiguous()):
        raise _lib.B2MError("copy_columns: column range outside the tensors, or tensors not contiguous")
    args = (src.data_ptr() + 2 * src_col, src.shape[1], dst.data_ptr() + 2 * dst_col, dst.shape[1], n, width)
    ll = LaunchList.current
    if ll is not None:
        Profile.launches += 1
        ll.add("copy_columns", ll.OP_COPY_COLUMNS, args)
    else:
        _run("copy_columns", 1, lambda: check(lib.b2m_copy_columns(*args, stream_ptr()), "copy_columns"),
             nbytes=4.0 * n * width)
    return dst


# ---------------------------------------------------------------------------------------------
# batch norm
# ---------------------------------------------------------------------------------------------
def colstats(x):
    lib = _lib_or_raise()
    _cuda(x, torch.bfloat16, "x")
    sums = ZeroArena.take(2 * x.shape[1], x.device)
    _run("colstats", 1, lambda: check(lib.b2m_colstats(ptr(x), x.shape[0], x.shape[1], ptr(sums), stream_ptr()),
                                      "colstats"), nbytes=2 * x.numel())
    return sums


def bn_forward(x, sums, gamma, beta, running_mean, running_var, momentum, eps, training, residual=None, relu=False,
               n_stat=None, want_mask=False):
    """want_mask (with relu): also returns the ReLU gate as uint8[n, c/8] bits for bn_backward(relu_mask=...)."""
    lib = _lib_or_raise()
    _cuda(x, torch.bfloat16, "x")
    n, c = x.shape
    out = torch.empty_like(x)
    save_mean = torch.empty(c, dtype=torch.float32, device=x.device)
    save_invstd = torch.empty(c, dtype=torch.float32, device=x.device)
    mask = torch.empty((n, c // 8), dtype=torch.uint8, device=x.device) if (want_mask and relu) else None
    ll = LaunchList.current
    if ll is not None:
        Profile.launches += 1
        ll.add("bn_forward", ll.OP_BN_FORWARD, (
            x.data_ptr(), n, n if n_stat is None else int(n_stat), c, _p(sums), gamma.data_ptr(), beta.data_ptr(),
            _p(running_mean), _p(running_var), 1 if training else 0, _p(residual), 1 if relu else 0, out.data_ptr(),
            save_mean.data_ptr(), save_invstd.data_ptr(), _p(mask)), float(momentum), float(eps))
        return (out, save_mean, save_invstd, mask) if want_mask else (out, save_mean, save_invstd)
    _run("bn_forward", 1, lambda: check(lib.b2m_bn_forward(
        ptr(x), n, n if n_stat is None else int(n_stat), c, ptr(sums), ptr(gamma), ptr(beta), ptr(running_mean),
        ptr(running_var), float(momentum), float(eps), int(bool(training)), ptr(residual), int(bool(relu)), ptr(out),
        ptr(save_mean), ptr(save_invstd), ptr(mask), stream_ptr()), "bn_forward"),
        nbytes=2 * x.numel() * (3 if residual is not None else 2))
    if want_mask:
        return out, save_mean, save_invstd, mask
    return out, save_mean, save_invstd


def bn_backward(x, out, dout, save_mean, save_invstd, gamma, relu, training, want_dresidual, n_stat=None,
                reduce_hook=None, n_stat_dev=None, dgamma=None, dbeta=None, relu_mask=None, red=None):
    """n_stat_dev: optional f64[1] device tensor with the global row count (SyncBN, no host round trip).
    reduce_hook(red) -> all-reduced copy of red: only dx uses it; dgamma / dbeta stay this rank's own sums.
    red: the reduction (sum g, sum g*xhat) f64[2c] if a dgrad epilogue already produced it (conv_forward(bn_reduce=...))."""
    lib = _lib_or_raise()
    n, c = x.shape
    have_red = red is not None
    if not have_red:
        red = ZeroArena.take(2 * c, x.device)
    ll = LaunchList.current
    if ll is not None:
        Profile.launches += 1 if have_red else 2
        rl, tr = 1 if relu else 0, 1 if training else 0
        if not have_red:
            ll.add("bn_backward_reduce", ll.OP_BN_BACKWARD_REDUCE, (
                x.data_ptr(), _p(out), dout.data_ptr(), n, c, save_mean.data_ptr(), save_invstd.data_ptr(), rl,
                red.data_ptr(), _p(relu_mask)))
        red_local = None
        if reduce_hook is not None:
            if not getattr(reduce_hook, "deferred", False):
                ll.flush()                  # the hook is an eager collective on the reduction
            red_local = red
            red = reduce_hook(red)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if want_dresidual else None
        if dgamma is None:
            dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
        if dbeta is None:
            dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
        ll.add("bn_backward_apply", ll.OP_BN_BACKWARD_APPLY, (
            x.data_ptr(), _p(out), dout.data_ptr(), n, n if n_stat is None else int(n_stat), c, save_mean.data_ptr(),
            save_invstd.data_ptr(), gamma.data_ptr(), red.data_ptr(), _p(red_local), _p(n_stat_dev), rl, tr, dx.data_ptr(),
            _p(dres), dgamma.data_ptr(), dbeta.data_ptr(), _p(relu_mask)))
        return dx, dres, dgamma, dbeta
    if not have_red:
        _run("bn_backward_reduce", 1, lambda: check(lib.b2m_bn_backward_reduce(
            ptr(x), ptr(out), ptr(dout), n, c, ptr(save_mean), ptr(save_invstd), int(bool(relu)), ptr(red), ptr(relu_mask),
            stream_ptr()), "bn_backward_reduce"), nbytes=2 * x.numel() * (3 if (relu and relu_mask is None) else 2))
    red_local = None
    if reduce_hook is not None:
        red_local = red
        red = reduce_hook(red)  # SyncBN: (sum_g, sum_g*xhat) summed over ranks, in a NEW buffer
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dresidual else None
    if dgamma is None:
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
    if dbeta is None:
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    _run("bn_backward_apply", 1, lambda: check(lib.b2m_bn_backward_apply(
        ptr(x), ptr(out), ptr(dout), n, n if n_stat is None else int(n_stat), c, ptr(save_mean), ptr(save_invstd),
        ptr(gamma), ptr(red), ptr(red_local), ptr(n_stat_dev), int(bool(relu)), int(bool(training)), ptr(dx), ptr(dres),
        ptr(dgamma), ptr(dbeta), ptr(relu_mask), stream_ptr()), "bn_backward_apply"),
        nbytes=2 * x.numel() * ((3 if (relu and relu_mask is None) else 2) + (2 if want_dresidual else 1)))
    return dx, dres, dgamma, dbeta


# ---------------------------------------------------------------------------------------------
# pooling
# ---------------------------------------------------------------------------------------------
def segment_mean_forward(f, ids, s):
    lib = _lib_or_raise()
    _cuda(f, torch.bfloat16, "f")
    _cuda(ids, torch.int64, "ids")
    out = torch.empty((s, f.shape[1]), dtype=torch.float32, device=f.device)
    counts = torch.empty(s, dtype=torch.float32, device=f.device)
    _run("segment_mean_forward", 2, lambda: check(lib.b2m_segment_mean_forward(
        ptr(f), ptr(ids), f.shape[0], f.shape[1], s, ptr(out), ptr(counts), stream_ptr()), "segment_mean_forward"),
        nbytes=2 * f.numel() + 8 * f.shape[0] + 4 * s * f.shape[1])
    return out, counts


def segment_mean_backward(dout, ids, counts, n):
    lib = _lib_or_raise()
    _cuda(dout, torch.float32, "dout")
    df = torch.empty((n, dout.shape[1]), dtype=torch.bfloat16, device=dout.device)
    _run("segment_mean_backward", 1, lambda: check(lib.b2m_segment_mean_backward(
        ptr(dout), ptr(ids), ptr(counts), n, dout.shape[1], dout.shape[0], ptr(df), stream_ptr()), "segment_mean_backward"),
        nbytes=2 * n * dout.shape[1] + 8 * n + 4 * dout.numel())
    return df


def segment_max_forward(f, ids, s):
    lib = _lib_or_raise()
    _cuda(f, torch.bfloat16, "f")
    out = torch.empty((s, f.shape[1]), dtype=torch.float32, device=f.device)
    argmax = torch.empty((s, f.shape[1]), dtype=torch.int32, device=f.device)
    check(lib.b2m_segment_max_forward(ptr(f), ptr(ids), f.shape[0], f.shape[1], s, ptr(out), ptr(argmax),
                                      stream_ptr()), "segment_max_forward")
    return out, argmax


def segment_max_backward(dout, argmax, n):
    lib = _lib_or_raise()
    _cuda(dout, torch.float32, "dout")
    _cuda(argmax, torch.int32, "argmax")
    s, c = dout.shape
    df = torch.empty((n, c), dtype=torch.bfloat16, device=dout.device)
    _run("segment_max_backward", 2, lambda: check(lib.b2m_segment_max_backward(
        ptr(dout), ptr(argmax), s, c, n, ptr(df), stream_ptr()), "segment_max_backward"), nbytes=2 * n * c + 8 * s * c)
    return df


# ---------------------------------------------------------------------------------------------
# decode
# ---------------------------------------------------------------------------------------------
def aabb_nms(boxes, cluster_th, max_clusters=None, want_heatmaps=True):
    """-> (representatives i64[K], cluster_of i32[M], heatmaps f32[K,M] or None). One host sync (K): the heat-map rows
    are allocated for the K clusters found (not for M) and filled by a second call."""
    lib = _lib_or_raise()
    _cuda(boxes, torch.float32, "boxes")
    m = boxes.shape[0]
    dev = boxes.device
    n_clusters = torch.zeros(1, dtype=torch.int32, device=dev)
    reps = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    cluster_of = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    ws_bytes = lib.b2m_nms_workspace_bytes(m)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _run("aabb_nms", 4, lambda: check(lib.b2m_aabb_nms(ptr(boxes), m, float(cluster_th), ptr(n_clusters), ptr(reps),
                                                       ptr(cluster_of), None, 0, ptr(ws), ws_bytes, stream_ptr()), "aabb_nms"),
         nbytes=28 * m + m * ((m + 31) // 32) * 4)
    k = int(n_clusters.item())
    heat = None
    if want_heatmaps:
        rows = k if max_clusters is None else min(k, int(max_clusters))
        heat = torch.empty((rows, m), dtype=torch.float32, device=dev)
        _run("aabb_heatmaps", 1, lambda: check(lib.b2m_aabb_heatmaps(ptr(boxes), m, ptr(n_clusters), ptr(reps), rows, ptr(heat),
                                                                    stream_ptr()), "aabb_heatmaps"), nbytes=4 * rows * m)
    return reps[:k].long(), cluster_of[:m], heat


def mask_label_vote(masks, label, n_vox, n_labels):
    """Per-mask majority label over bit-packed masks int32[k, words]; label int32[n_vox] -> (best i32[k], counts i32[k, L])."""
    lib = _lib_or_raise()
    _cuda(masks, torch.int32, "masks")
    _cuda(label, torch.int32, "label")
    k, words = masks.shape
    counts = torch.empty((k, n_labels), dtype=torch.int32, device=masks.device)
    best = torch.empty(k, dtype=torch.int32, device=masks.device)
    _run("mask_label_vote", 3, lambda: check(lib.b2m_mask_label_vote(ptr(masks), k, words, n_vox, ptr(label), n_labels,
                                                                    ptr(counts), ptr(best), stream_ptr()), "mask_label_vote"),
         nbytes=4 * k * words + 4 * n_vox)
    return best, counts


def segment_label_vote(seg, label, n_seg, n_labels):
    """Mode of label int32[n] within each segment id seg int64[n] -> (best i32[n_seg], counts i32[n_seg, L])."""
    lib = _lib_or_raise()
    _cuda(seg, torch.int64, "seg")
    _cuda(label, torch.int32, "label")
    n = seg.shape[0]
    counts = torch.empty((n_seg, n_labels), dtype=torch.int32, device=seg.device)
    best = torch.empty(n_seg, dtype=torch.int32, device=seg.device)
    _run("segment_label_vote", 3, lambda: check(lib.b2m_segment_label_vote(ptr(seg), ptr(label), n, n_seg, n_labels, ptr(counts),
                                                                          ptr(best), stream_ptr()), "segment_label_vote"),
         nbytes=12 * n)
    return best, counts


def heatmap_project(heat, fg_rank, seg2vox, mask_bin_th):
    lib = _lib_or_raise()
    _cuda(heat, torch.float32, "heat")
    k, m_fg = heat.shape
    n_vox = seg2vox.shape[0]
    words = (n_vox + 31) // 32
    masks = torch.empty((k, words), dtype=torch.int32, device=heat.device)
    check(lib.b2m_heatmap_project(ptr(heat), k, m_fg, ptr(fg_rank), ptr(seg2vox), n_vox, float(mask_bin_th), ptr(masks),
                                  stream_ptr()), "heatmap_project")
    return masks


def mask_nms(masks, th):
    """bit-packed masks int32[k, words] in score order -> keep bool[k]."""
    lib = _lib_or_raise()
    k, words = masks.shape
    keep = torch.empty(max(k, 1), dtype=torch.uint8, device=masks.device)
    n_keep = torch.zeros(1, dtype=torch.int32, device=masks.device)
    ws_bytes = lib.b2m_mask_nms_workspace_bytes(k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=masks.device)
    check(lib.b2m_mask_nms(ptr(masks), k, words, float(th), ptr(keep), ptr(n_keep), ptr(ws), ws_bytes, stream_ptr()),
          "mask_nms")
    return keep[:k].bool()


def unpack_masks(masks, n_vox):
    lib = _lib_or_raise()
    k, words = masks.shape
    out = torch.empty((k, n_vox), dtype=torch.uint8, device=masks.device)
    check(lib.b2m_unpack_masks(ptr(masks), k, words, n_vox, ptr(out), stream_ptr()), "unpack_masks")
    return out.bool()


def pack_masks_torch(masks_bool):
    """bool[k, n] -> int32[k, ceil(n/32)] bit-packed (host-side helper for tests; little-endian bits)."""
    k, n = masks_bool.shape
    words = (n + 31) // 32
    pad = words * 32 - n
    m = torch.nn.functional.pad(masks_bool.to(torch.int64), (0, pad)).view(k, words, 32)
    weights = (2 ** torch.arange(32, dtype=torch.int64, device=masks_bool.device))
    w = (m * weights).sum(-1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.to(torch.int32).contiguous()


# ---------------------------------------------------------------------------------------------
# voxelisation (the step before the path; models/dataloader.py:61-77)
# ---------------------------------------------------------------------------------------------
def voxel_coords(positions, min_position, voxel_size):
    """positions f64[P,3], min_position f64[1] (device) -> int32[P,4] rounded voxel coordinate of every point, status."""
    lib = _lib_or_raise()
    _cuda(positions, torch.float64, "positions")
    _cuda(min_position, torch.float64, "min_position")
    p = positions.shape[0]
    coords = torch.empty((p, 4), dtype=torch.int32, device=positions.device)
    status = torch.zeros(1, dtype=torch.int32, device=positions.device)
    _run("voxel_coords", 1, lambda: check(lib.b2m_voxel_coords(
        ptr(positions), p, ptr(min_position), float(voxel_size), ptr(coords), ptr(status), stream_ptr()), "voxel_coords"),
        nbytes=24 * p + 16 * p)
    return coords, status


def nearest_point(positions, min_position, voxel_size, vox_coords, nbr, start, point_order):
    """index of the scene point nearest to every voxel centre, int64[n_vox] (see csrc/voxel.cu)."""
    lib = _lib_or_raise()
    _cuda(positions, torch.float64, "positions")
    _cuda(vox_coords, torch.int32, "vox_coords")
    _cuda(start, torch.int64, "start")
    _cuda(point_order, torch.int64, "point_order")
    n_vox = vox_coords.shape[0]
    out = torch.empty(n_vox, dtype=torch.int64, device=positions.device)
    _run("nearest_point", 1, lambda: check(lib.b2m_nearest_point(
        ptr(positions), ptr(min_position), float(voxel_size), ptr(vox_coords), n_vox, ptr(nbr), ptr(start),
        ptr(point_order), ptr(out), stream_ptr()), "nearest_point"), nbytes=24 * positions.shape[0] + 27 * 4 * n_vox)
    return out
