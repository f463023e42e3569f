"""The multi-GPU path on real devices: CudaBackend partials + one NCCL all_gather + rank-0 combine, checked against
the oracle.  Runs with as many ranks as there are GPUs (1 on the driver's single-GPU box: a 1-rank NCCL group still
goes through the same collective call)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    import zkvm_b200 as zk
    from oracle import c_oracle
    from zkvm_b200.sharded import CudaBackend, shard_range, sharded_msm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(99)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    lo, hi = shard_range(n, rank, world)
    ctx = zk.Context(rank)
    table = zk.PointTable(ctx, hi - lo).append_compressed(pts[32 * lo:32 * hi])        # each rank caches ITS slice
    backend = CudaBackend(ctx, table)
    local = torch.from_numpy(sc[lo:hi].copy()).reshape(-1).cuda(rank)
    out = sharded_msm(backend, local, n)
    if rank == 0:
        q.put((out, c_oracle.msm(sc, pts, n, threads=4)))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [5000, (1 << 16) + 3])
def test_sharded_msm_nccl(n):
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    assert world >= 1
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs: p.start()
    got, want = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want and got is not None


def test_single_process_no_group(ctx, c_oracle):
    import zkvm_b200 as zk
    from zkvm_b200.sharded import CudaBackend, sharded_msm
    n = 3000
    rng = np.random.default_rng(5)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    table = zk.PointTable(ctx, n).append_compressed(pts)
    out = sharded_msm(CudaBackend(ctx, table), torch.from_numpy(sc).reshape(-1).cuda(), n)
    assert out == c_oracle.msm(sc, pts, n)
