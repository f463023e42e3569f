"""Point-range sharding of one large MSM across the GPUs of a box (BASELINE.json config 5).

An MSM is a sum over independent terms, so rank r of G takes the contiguous index range
shard_range(n, r, G), computes one partial sum on its own GPU (no data-path collective), and the G
partial points -- 128 bytes each, extended coordinates -- are gathered with a single NCCL collective;
rank 0 adds them and encodes.  The gather is latency-bound (G*128 bytes), not bandwidth-bound.

Two front ends:
* `msm_single_process(...)`: ONE process drives all GPUs through the C ABI's zk_mgpu_* entry points (what a Rust
  verifier calls; partials gathered by peer copies or one ncclAllGather inside the library).  This module only
  forwards to it.
* `sharded_msm(...)`: one process per GPU under torch.distributed (how bench.py's scaling runs are launched); the
  compute backend is injected so that the host logic (partitioning, gather order, combine) can be tested with gloo
  on CPU; the product backend is `CudaBackend`, which goes through the C ABI.
"""
from __future__ import annotations

from typing import Optional, Protocol

import torch
import torch.distributed as dist

PARTIAL_BYTES = 128


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced, order-preserving partition of range(n): the first n % world shards get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Backend(Protocol):
    device: torch.device
    def partial(self, scalars: torch.Tensor, lo: int, hi: int) -> torch.Tensor: ...   # -> uint8[128] on self.device
    def combine(self, partials: torch.Tensor) -> bytes: ...                            # uint8[G,128] -> 32 bytes


class CudaBackend:
    """Each rank caches ITS slice of the points on its GPU once (`load_compressed` / `load_uniform`);
    scalars for that slice arrive per call as a CUDA uint8 tensor."""

    def __init__(self, ctx, table):
        self.ctx, self.table = ctx, table
        self.device = torch.device("cuda", ctx.device)
        self._out = torch.empty(PARTIAL_BYTES, dtype=torch.uint8, device=self.device)

    def partial(self, scalars: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
        n = hi - lo
        assert scalars.is_cuda and scalars.dtype == torch.uint8 and scalars.numel() == 32 * n and scalars.is_contiguous()
        # torch's current stream produced `scalars`; the ctx stream consumes them
        torch.cuda.current_stream(self.device).synchronize()
        self.ctx.msm_table_dev(scalars.data_ptr(), self.table, 0, n, self._out.data_ptr())
        self.ctx.sync()
        return self._out

    def combine(self, partials: torch.Tensor) -> bytes:
        assert partials.is_cuda and partials.is_contiguous()
        torch.cuda.current_stream(self.device).synchronize()
        return bytes(self.ctx.ext_sum_compress_dev(partials.data_ptr(), partials.shape[0]))


def sharded_msm(backend: Backend, local_scalars: torch.Tensor, n_total: int, group=None) -> Optional[bytes]:
    """Every rank calls this with the scalars of its own index range.  Returns the 32-byte encoding on rank 0,
    None elsewhere.  Exactly one collective: an all_gather of 128 bytes per rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n_total, rank, world)
    part = backend.partial(local_scalars, lo, hi)
    if world == 1:
        return backend.combine(part.view(1, PARTIAL_BYTES))
    gathered = torch.empty(world, PARTIAL_BYTES, dtype=torch.uint8, device=backend.device)
    dist.all_gather_into_tensor(gathered, part.view(1, PARTIAL_BYTES), group=group)
    return backend.combine(gathered) if rank == 0 else None


def msm_single_process(scalars, points, devices=None, g: Optional[int] = None, gather: str = "peer", mg=None) -> Optional[bytes]:
    """One process, several GPUs: forwards to zk_mgpu_msm_vartime (include/zkmsm.h).  `scalars` / `points` are host
    buffers of n x 32 bytes (compressed encodings).  Returns the 32-byte encoding, or None for an invalid encoding.
    Pass a long-lived `mg` (zkvm_b200.MultiGpu) to amortise context creation."""
    from .ristretto import MultiGpu
    own = mg is None
    if own:
        mg = MultiGpu(devices=devices, g=g, gather=gather)
    try:
        r = mg.optional_multiscalar_mul(scalars, points)
        return None if r is None else bytes(r)
    finally:
        if own:
            mg.close()
