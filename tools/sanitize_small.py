#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
from oracle import c_oracle
ctx = zk.Context(0)
rng = np.random.default_rng(3)
for n in (1, 37, 3001):
    u = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    tab = zk.PointTable(ctx, 1).append_uniform(u)
    pts = tab.compress()
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    want = c_oracle.msm(sc, pts, n)
    for c in (0, 4, 13, 16):
        ctx.set_window(c)
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want
    ctx.set_window(0)
    same = bytes(sc[0]) * n      # everything in one bucket per window: the warp-cooperative path
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, same, pts)) == c_oracle.msm(same, pts, n)
# round-2 kernels: batched normalisation (unchecked extended ingestion), validated ingestion, one-launch point sum,
# window expansion, the multi-GPU layer (peer-copy gather) and the staging ring
from oracle import ristretto255_ref as ref
blob = b"".join(b"".join((v * 7 % ref.P).to_bytes(32, "little") for v in (p.X, p.Y, p.Z, p.T))
                for p in (ref.decode(pts[32 * i:32 * i + 32]) for i in range(0, 300, 10)))
want_pts = b"".join(pts[32 * i:32 * i + 32] for i in range(0, 300, 10))
assert zk.PointTable(ctx).append_extended(blob).compress() == want_pts
assert zk.PointTable(ctx).append_extended_unchecked(blob).compress() == want_pts
assert bytes(ctx.sum_compressed(pts[:32 * 300])) == c_oracle.point_sum(pts[:32 * 300], 300)
tab.precompute(9)
assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want
ctx.set_staging(2)
assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want
ctx.set_staging(0)
mg = zk.MultiGpu(g=1)
assert bytes(mg.optional_multiscalar_mul(sc, pts)) == want
mt = zk.MultiGpuTable(mg).append_compressed(pts)
assert bytes(mg.vartime_multiscalar_mul(sc, mt)) == want
print("sanitize run ok")
