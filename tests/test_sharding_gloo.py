"""Host-side logic of the multi-GPU path with world_size 2 over gloo on CPU: index-range partition, the single
gather, rank-0 combine.  The compute backend is the oracle here (this is a test); on GPUs it is CudaBackend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zkvm_b200.sharded import PARTIAL_BYTES, shard_range, sharded_msm


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 9, 1 << 20, (1 << 22) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class OracleBackend:
    """partial = 32-byte oracle encoding padded to the 128-byte partial slot; combine = oracle point sum."""
    device = torch.device("cpu")

    def __init__(self, points):
        self.points = points

    def partial(self, scalars, lo, hi):
        from oracle import c_oracle
        enc = c_oracle.msm(scalars.numpy(), self.points[32 * lo:32 * hi], hi - lo)
        return torch.frombuffer(bytearray(enc + bytes(PARTIAL_BYTES - 32)), dtype=torch.uint8)

    def combine(self, partials):
        from oracle import c_oracle
        g = partials.shape[0]
        return c_oracle.point_sum(b"".join(bytes(partials[i, :32].tolist()) for i in range(g)), g)


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    rng = np.random.default_rng(1234)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    lo, hi = shard_range(n, rank, world)
    out = sharded_msm(OracleBackend(pts), torch.from_numpy(sc[lo:hi].copy()).reshape(-1), n)
    if rank == 0:
        q.put((out, c_oracle.msm(sc, pts, n)))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4096])
def test_world2_gloo(n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs: p.start()
    got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == want and got is not None
