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
