#!/usr/bin/env python3
"""Minimal driver for ncu: a few n-point MSMs over a cached table (and one from compressed input)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << logn
ctx = zk.Context(0)
rng = np.random.default_rng(1)
tab = zk.PointTable(ctx, n).append_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8))
comp = tab.compress()
sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
for _ in range(reps):
    r = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
r2 = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, comp)
assert bytes(r) == bytes(r2)
print("ok", bytes(r).hex())
