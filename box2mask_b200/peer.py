"""SyncBatchNorm statistics exchanged over NVLink peer memory (include/b2m.h, "SyncBatchNorm statistics over NVLink peer
memory"; csrc/peer.cu).

The reference's multi-GPU configuration turns every BatchNorm into SyncBatchNorm (/root/reference/models/model.py:25):
162 tiny dependent all-reduces per training step. On one node with one process per GPU `PeerExchange` replaces them by a
single-CTA kernel per exchange that writes the rank's sums straight into its peers' memory; the kernel is stream-ordered
and a launch-list command, so the pass is no longer drained once per layer. Multi-node groups, CPU tensors (the gloo
tests) and vectors longer than the exchange slot keep using torch.distributed - a choice of transport by topology, the
arithmetic (a sum in rank order) is the same.
"""
import ctypes
import socket

import torch

from . import _lib, ops


class PeerExchange:
    _by_group = {}      # (id(group), device) -> PeerExchange or None (None = group not eligible: use the library collective)
    CHECK_EVERY = 4096  # exchanges between two reads of the time-out flag

    @classmethod
    def for_group(cls, group, device):
        """The exchange of a process group whose ranks all sit on this node, else None. Collective: every rank of the
        group must call it at the same point the first time."""
        device = torch.device(device)
        if device.type != "cuda" or group is None:
            return None
        key = (id(group), device)
        if key not in cls._by_group:
            cls._by_group[key] = cls._create(group, device)
        return cls._by_group[key]

    @classmethod
    def _create(cls, group, device):
        import torch.distributed as dist
        ops.LaunchList.flush_current()          # set-up runs eager collectives and copies
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > 16 or dist.get_backend(group) != "nccl":
            return None
        lib = _lib.load()
        buf, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        ok = True
        with torch.cuda.device(device):
            ok = lib.b2m_peer_buffer_create(ctypes.byref(buf), handle) == 0
        infos = [None] * world
        dist.all_gather_object(infos, (socket.gethostname(), bytes(handle.raw) if ok else None), group=group)
        ok = ok and all(h == infos[0][0] and raw is not None for h, raw in infos)
        peers = [None] * world
        if ok:
            with torch.cuda.device(device):
                for r, (_, raw) in enumerate(infos):
                    if r == rank:
                        peers[r] = buf.value
                        continue
                    p = ctypes.c_void_p()
                    if lib.b2m_peer_buffer_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(p)) != 0:
                        ok = False
                        break
                    peers[r] = p.value
        # every rank must take the same path: one agreement round
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            with torch.cuda.device(device):
                for r, p in enumerate(peers):
                    if p is not None and r != rank:
                        lib.b2m_peer_buffer_close(ctypes.c_void_p(p), 0)
                if buf.value:
                    lib.b2m_peer_buffer_close(buf, 1)
            return None
        return cls(device, rank, world, peers)

    def __init__(self, device, rank, world, peers):
        self.device, self.rank, self.world = device, rank, world
        self.peers = peers                      # raw device pointers, kept mapped for the life of the process
        self.table = torch.tensor(peers, dtype=torch.int64, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.seq = 0
        self.max_doubles = int(_lib.load().b2m_peer_max_doubles())

    def fits(self, n):
        return n <= self.max_doubles

    def allreduce(self, vec, tail=None):
        """-> new f64 tensor: the sum of `vec` over the ranks (element -1 of each rank's vector replaced by `tail` when
        given - the row count that travels with the BatchNorm sums). Stream-ordered; no host synchronisation."""
        if vec.dtype != torch.float64 or not vec.is_contiguous() or vec.device != self.device:
            raise _lib.B2MError("peer exchange: contiguous float64 vector on %s expected" % self.device)
        n = vec.numel()
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        self.seq += 1
        args = (vec.data_ptr(), n, 1 if tail is not None else 0, out.data_ptr(), self.table.data_ptr(), self.rank, self.world,
                self.seq, self.status.data_ptr())
        t = float(tail) if tail is not None else 0.0
        ll = ops.LaunchList.current
        ops.Profile.launches += 1
        if ll is not None:
            ll.add("peer_allreduce", ll.OP_PEER_ALLREDUCE, args, t)
        else:
            _lib.check(_lib.load().b2m_peer_allreduce_f64(args[0], n, t, args[2], args[3], args[4], self.rank, self.world,
                                                           self.seq, args[8], _lib.stream_ptr(self.device)), "peer_allreduce")
        if self.seq % self.CHECK_EVERY == 0:
            self.check()
        return out

    def check(self):
        """Raises if an exchange timed out (a peer never arrived). Reads a flag back from the device."""
        ops.LaunchList.flush_current()
        if int(self.status.item()) != 0:
            raise _lib.B2MError("SyncBatchNorm peer exchange timed out on rank %d: a peer did not arrive" % self.rank)


def allreduce_sum(vec, group, tail=None):
    """Sum of a float64 vector over `group`: over NVLink peer memory when the group qualifies (PeerExchange), through
    torch.distributed otherwise. tail: see PeerExchange.allreduce (the vector's last element is then a placeholder)."""
    px = PeerExchange.for_group(group, vec.device) if vec.is_cuda else None
    if px is not None and px.fits(vec.numel()) and vec.dtype == torch.float64:
        return px.allreduce(vec.contiguous(), tail)
    ops.LaunchList.flush_current()          # an eager collective on a vector that pending launches produce
    out = vec.clone()
    if tail is not None:
        out[-1] = float(tail)
    torch.distributed.all_reduce(out, group=group)
    return out
