#!/usr/bin/env python3
"""Randomised stress: every entry point against the C oracle for `seconds` seconds.  usage: stress.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
from oracle import c_oracle
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
ctx = zk.Context(0)
L = zk.GROUP_ORDER
pool_n = 1 << 17
pool = c_oracle.from_uniform(rng.integers(0, 256, size=(pool_n, 64), dtype=np.uint8), pool_n)
plain = zk.PointTable(ctx).append_compressed(pool)
import torch
mg = zk.MultiGpu(g=min(torch.cuda.device_count(), 8))
t_end = time.time() + seconds
it = 0
def scalars(n):
    kind = int(rng.integers(0, 6))
    if kind == 0: return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    if kind == 1: return np.tile(rng.integers(0, 256, size=(1, 32), dtype=np.uint8), (n, 1))
    if kind == 2:
        s = np.zeros((n, 32), dtype=np.uint8); s[:, int(rng.integers(0, 32))] = rng.integers(0, 256, size=n, dtype=np.uint8); return s
    if kind == 3:
        vals = [0, 1, 2, L - 1, L, L + 1, 2**252, 2**255 - 1, 2**256 - 1, 8 * L]
        return np.frombuffer(b"".join((vals[int(i)] % 2**256).to_bytes(32, "little") for i in rng.integers(0, len(vals), size=n)), dtype=np.uint8).reshape(n, 32)
    if kind == 4:
        s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s[:, 16:] = 0; return s          # 128-bit scalars
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s[rng.random(n) < 0.5] = 0; return s   # half zero
while time.time() < t_end:
    it += 1
    n = int(2 ** rng.uniform(0, 17)); n = min(n, pool_n)
    off = int(rng.integers(0, pool_n - n + 1))
    sc = scalars(n); pts = pool[32 * off:32 * (off + n)]
    want = c_oracle.msm(sc, pts, n, threads=8)
    ctx.set_window(int(rng.choice([0, 0, 0, 4, 6, 8, 10, 11, 12, 13, 14, 15, 16])))
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want, ("compressed", it, n)
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, plain, offset=off)) == want, ("table", it, n)
    k = int(rng.integers(0, n + 1))
    assert bytes(zk.RistrettoPoint.mixed_multiscalar_mul(ctx, sc[:k], plain, sc[k:], pts[32 * k:], offset=off)) == want, ("mixed", it, n, k)
    ctx.set_window(0)
    if it % 3 == 0:                                   # single-process multi-GPU entry point, pageable and staged sources
        mg.set_staging(2 if it % 6 == 0 else 0)
        assert bytes(mg.optional_multiscalar_mul(sc, pts)) == want, ("mgpu", it, n)
    if it % 4 == 0:
        ctx.set_staging(2)
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want, ("staged", it, n)
        ctx.set_staging(0)
    if it % 8 == 0:
        g = min(n, int(rng.integers(1, 1025)))
        assert bytes(ctx.sum_compressed(pts[:32 * g])) == c_oracle.point_sum(pts[:32 * g], g), ("sum", it, g)
    if it % 7 == 0:
        pre = zk.PointTable(ctx).append_compressed(pts).precompute(int(rng.integers(4, 21)))
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, pre)) == want, ("precomputed", it, n)
        pre.close()
    if it % 5 == 0:
        m = int(rng.integers(1, 40))
        cuts = np.sort(rng.integers(0, n + 1, size=m - 1)) if m > 1 else np.array([], dtype=np.int64)
        seg = np.concatenate([[0], cuts, [n]]).astype(np.uint64)
        got = zk.batch_optional_multiscalar_mul(ctx, sc, pts, seg)
        for j, (a, b) in enumerate(zip(seg, seg[1:])):
            a, b = int(a), int(b)
            assert bytes(got[j]) == c_oracle.msm(sc[a:b], pts[32 * a:32 * b], b - a), ("batch", it, n, j)
    if it % 11 == 0:
        # m MSMs over one shared (sometimes window-expanded) table
        m = int(rng.integers(1, 12)); per = int(rng.integers(1, max(2, n // m + 1)))
        per = min(per, n)
        seg = np.arange(0, m * per + 1, per, dtype=np.uint64)
        bs = scalars(m * per)
        shared = zk.PointTable(ctx).append_compressed(pts[:32 * per])
        if it % 22 == 0: shared.precompute(int(rng.integers(4, 21)))
        got = zk.batch_vartime_multiscalar_mul(ctx, bs, shared, seg)
        for j in range(m):
            assert bytes(got[j]) == c_oracle.msm(bs[j * per:(j + 1) * per], pts[:32 * per], per), ("table batch", it, j)
        shared.close()
print(f"stress ok: {it} iterations in {seconds:.0f} s, seed {seed}")
