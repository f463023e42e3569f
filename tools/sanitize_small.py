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
print("sanitize run ok")
