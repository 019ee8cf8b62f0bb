"""2 GPUs, NCCL: the bf16 SyncBatchNorm path (functional.BatchNormFn with a process group) against a single-process run
on the concatenated batch — outputs, input gradients and the AFFINE gradients, which must stay rank-local sums (torch's
SyncBatchNorm semantics; the gradient all-reduce averages them afterwards). Skipped on a 1-GPU box (run it with
`gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda:%d" % rank))
    try:
        from box2mask_b200 import functional as Fn
        dev = "cuda:%d" % rank
        g = torch.Generator().manual_seed(0)
        n, c = 5000, 64
        x_full = torch.randn(n, c, generator=g).to(torch.bfloat16)
        res_full = torch.randn(n, c, generator=g).to(torch.bfloat16)
        go_full = torch.randn(n, c, generator=g).to(torch.bfloat16)
        gamma0 = torch.rand(c, generator=g) + 0.5
        beta0 = torch.randn(c, generator=g) * 0.1
        cut = 1800                                  # ranks hold different row counts
        sl = slice(0, cut) if rank == 0 else slice(cut, n)

        def run(x, res, go, group):
            x = x.to(dev).requires_grad_(True)
            res = res.to(dev).requires_grad_(True)
            gamma = gamma0.clone().to(dev).requires_grad_(True)
            beta = beta0.clone().to(dev).requires_grad_(True)
            rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
            y = Fn.BatchNormFn.apply(x, None, gamma, beta, rm, rv, 0.1, 1e-5, True, res, True, group)
            y.backward(go.to(dev))
            return y.detach().float(), x.grad.float(), res.grad.float(), gamma.grad, beta.grad, rm, rv
        y, dx, dres, dg, db, rm, rv = run(x_full[sl], res_full[sl], go_full[sl], dist.group.WORLD)
        yf, dxf, dresf, dgf, dbf, rmf, rvf = run(x_full, res_full, go_full, None)
        assert torch.allclose(y, yf[sl], atol=2e-2, rtol=2e-2), float((y - yf[sl]).abs().max())
        assert torch.allclose(dx, dxf[sl], atol=2e-2, rtol=2e-2), float((dx - dxf[sl]).abs().max())
        assert torch.equal(dres, dresf[sl])
        assert torch.allclose(rm, rmf, atol=1e-5) and torch.allclose(rv, rvf, atol=1e-5)
        # affine gradients are LOCAL sums: summed over ranks they equal the full-batch gradients (an average over ranks
        # times the world size, which is what the gradient all-reduce + a world-scaled loss produce)
        tot_g, tot_b = dg.clone(), db.clone()
        dist.all_reduce(tot_g)
        dist.all_reduce(tot_b)
        assert torch.allclose(tot_g, dgf, rtol=1e-3, atol=1e-2), float((tot_g - dgf).abs().max())
        assert torch.allclose(tot_b, dbf, rtol=1e-3, atol=1e-2), float((tot_b - dbf).abs().max())
        assert float((dg - dgf).abs().max()) > 1e-3          # and they are NOT already the global sums
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sync_batchnorm_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}


def _peer_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda:%d" % rank))
    try:
        from box2mask_b200 import ops
        from box2mask_b200.peer import PeerExchange, allreduce_sum
        dev = torch.device("cuda:%d" % rank)
        px = PeerExchange.for_group(dist.group.WORLD, dev)
        assert px is not None, "ranks of one node must get the NVLink peer exchange"
        g = torch.Generator().manual_seed(100 + rank)
        # 300 exchanges of different lengths back to back (the two slot sets alternate, a fast rank runs ahead), every
        # one compared with the NCCL all-reduce of the same vector; with and without the row-count tail
        sizes = [1, 2, 129, 513, 1025, 2048]
        outs, refs = [], []
        for i in range(300):
            n = sizes[i % len(sizes)]
            v = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
            tail = float(1000 * rank + i) if i % 3 == 0 else None
            ref = v.clone()
            if tail is not None:
                ref[-1] = tail
            dist.all_reduce(ref)
            outs.append(px.allreduce(v, tail))
            refs.append(ref)
            if rank == 1 and i % 50 == 0:
                torch.cuda._sleep(20_000_000)          # let the other rank run ahead
        for o, r in zip(outs, refs):
            assert torch.allclose(o, r, rtol=1e-14, atol=1e-14), float((o - r).abs().max())
        # the sums are bit-identical on all ranks (slots added in rank order)
        mine = torch.cat(outs)
        other = mine.clone()
        dist.broadcast(other, src=0)
        assert torch.equal(mine, other)
        # as launch-list commands, interleaved with the producing kernels of a pass
        x = (torch.randn(4096, 64, generator=g) + rank).to(dev).to(torch.bfloat16)
        sums = ops.colstats(x)
        packed = torch.cat([sums, torch.zeros(1, dtype=torch.float64, device=dev)])
        ll = ops.LaunchList.begin(dev)
        tot = allreduce_sum(packed, dist.group.WORLD, tail=4096.0)
        ll.end()
        ref = packed.clone()
        ref[-1] = 4096.0
        dist.all_reduce(ref)
        assert torch.allclose(tot, ref, rtol=1e-14, atol=1e-12) and float(tot[-1]) == 4096.0 * world
        # vectors longer than a slot fall back to the library collective
        big = torch.ones(5000, dtype=torch.float64, device=dev)
        assert float(allreduce_sum(big, dist.group.WORLD).sum()) == 5000.0 * world
        px.check()
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_peer_exchange_two_gpus():
    """b2m_peer_allreduce_f64 (SyncBatchNorm statistics over NVLink peer memory) against NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)          # every GPU of the box: 2 under `gpurun --gpus 2`, 8 on a full node
    out = mp.Manager().dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}
