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
        gs, n = Fn.sync_bn_stats(sums, mine.shape[0], dist.group.WORLD)
        assert n == 50
        assert torch.allclose(gs, torch.cat([full.sum(0), (full * full).sum(0)]))
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
