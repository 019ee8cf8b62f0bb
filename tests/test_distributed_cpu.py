"""CPU, world_size 2, gloo: the host-side multi-GPU logic (SyncBN statistics exchange, DDP gradient averaging over
the reference-named modules, per-rank scene sharding). The CUDA kernels themselves are rank-local."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import box2mask_b200
        ME = box2mask_b200.install_as_minkowski_engine()
        from box2mask_b200 import functional as Fn
        from box2mask_b200.me.sparse_tensor import CoordinateManager, SparseTensor
        torch.manual_seed(0)
        # 1. statistics exchange: ranks hold different row counts
        full = torch.randn(50, 8, dtype=torch.float64)
        mine = full[:20] if rank == 0 else full[20:]
        sums = torch.cat([mine.sum(0), (mine * mine).sum(0)])
        packed = Fn.sync_bn_stats(sums, mine.shape[0], dist.group.WORLD)    # [2C sums | global row count], on the device
        assert packed.shape == (17,) and float(packed[-1]) == 50
        assert torch.allclose(packed[:-1], torch.cat([full.sum(0), (full * full).sum(0)]))
        # 1b. the transport choice behind it (box2mask_b200/peer.py): CPU tensors / gloo groups never get the NVLink peer
        # exchange, allreduce_sum falls back to the library collective with the same semantics (tail = the row count
        # that replaces the vector's last element on every rank)
        from box2mask_b200.peer import PeerExchange, allreduce_sum
        assert PeerExchange.for_group(dist.group.WORLD, "cpu") is None
        vec = torch.arange(5, dtype=torch.float64) * (rank + 1)
        tot = allreduce_sum(vec, dist.group.WORLD, tail=10.0 + rank)
        assert torch.equal(tot, torch.tensor([0.0, 3.0, 6.0, 9.0, 21.0], dtype=torch.float64)) and float(vec[-1]) == 4.0 * (rank + 1)
        assert torch.equal(allreduce_sum(vec, dist.group.WORLD), torch.arange(5, dtype=torch.float64) * 3)
        # 2. an MLP head (reference naming) under DDP + SyncBN == the same head on the concatenated batch
        def head():
            torch.manual_seed(1)
            return torch.nn.Sequential(
                ME.MinkowskiConvolution(8, 8, kernel_size=1, bias=True, dimension=3), ME.MinkowskiReLU(),
                ME.MinkowskiBatchNorm(8), ME.MinkowskiConvolution(8, 3, kernel_size=1, bias=True, dimension=3))
        net = head()
        ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(net)
        assert isinstance(net[2], ME.MinkowskiSyncBatchNorm)
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        x_full = torch.randn(50, 8)
        x = x_full[:20] if rank == 0 else x_full[20:]
        cm = CoordinateManager(torch.zeros((x.shape[0], 4), dtype=torch.int32))
        y = ddp(SparseTensor(x, coordinate_manager=cm)).F
        # DDP averages gradients over ranks; weighting the local loss by world/size makes the sum the global mean loss
        loss = (y ** 2).sum() / 50 * world
        loss.backward()
        ref = head()
        cm2 = CoordinateManager(torch.zeros((50, 4), dtype=torch.int32))
        yr = ref(SparseTensor(x_full, coordinate_manager=cm2)).F
        ((yr ** 2).sum() / 50).backward()
        sl = slice(0, 20) if rank == 0 else slice(20, 50)
        assert torch.allclose(y, yr[sl], atol=1e-5), float((y - yr[sl]).abs().max())
        for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p.grad, q.grad, atol=1e-5), (k, float((p.grad - q.grad).abs().max()))
        assert torch.allclose(net[2].bn.running_mean, ref[2].bn.running_mean, atol=1e-6)
        assert torch.allclose(net[2].bn.running_var, ref[2].bn.running_var, atol=1e-5)
        # 2b. the same head under FlatGradSync (one flat all-reduce at the end of backward, grad_sync.py): identical
        #     gradients, p.grad re-pointed at slices of the flat buffer, parameters broadcast from rank 0 at construction
        from box2mask_b200.grad_sync import FlatGradSync
        net2 = head()
        ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(net2)
        if rank == 1:
            with torch.no_grad():
                for p in net2.parameters():
                    p.add_(1.0)               # must be overwritten by rank 0's values
        gs2 = FlatGradSync(net2)
        for step in range(2):                 # second pass: set_to_none=False keeps the views and accumulates in place
            net2.zero_grad(set_to_none=(step == 0))
            cm3 = CoordinateManager(torch.zeros((x.shape[0], 4), dtype=torch.int32))
            y2 = net2(SparseTensor(x, coordinate_manager=cm3)).F
            ((y2 ** 2).sum() / 50 * world).backward()
            assert gs2.syncs == step + 1
            for (k, p), (_, q) in zip(net2.named_parameters(), ref.named_parameters()):
                assert torch.allclose(p.grad, q.grad, atol=1e-5), (k, float((p.grad - q.grad).abs().max()))
            assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(gs2.params, gs2.views))
        # 3. per-rank scene sharding of the bench: different scenes per rank, deterministic per rank
        import bench
        a = bench.make_scenes(1, seed=10 + rank, scale=0.1)[0]["vox_coords"]
        b = bench.make_scenes(1, seed=10 + rank, scale=0.1)[0]["vox_coords"]
        assert np.array_equal(a, b)
        sizes = [None, None]
        dist.all_gather_object(sizes, int(a.sum()))
        assert sizes[0] != sizes[1]
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}
