#!/usr/bin/env python3
"""Empirical window sweep: for each n, time the table-path MSM at every window width (device phases, ms)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
ctx = zk.Context(0); ctx.set_profiling(True)
rng = np.random.default_rng(1)
for logn in [int(a) for a in sys.argv[1:]] or [10, 12, 14, 16, 18, 20]:
    n = 1 << logn
    tab = zk.PointTable(ctx, n).append_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8))
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    row = {}
    for c in range(4, 17):
        ctx.set_window(c)
        acc = np.zeros(4)
        for i in range(5):
            zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
            if i >= 2: acc += np.array(ctx.last_phase_ms())
        acc /= 3
        row[c] = [round(float(x), 3) for x in acc[1:]] + [round(float(acc[1:].sum()), 3)]
    best = min(row, key=lambda c: row[c][3])
    print(json.dumps({"logn": logn, "best_c": best, "best_ms": row[best][3], "picked": zk.pick_window(n), "picked_ms": row[zk.pick_window(n)][3],
                      "all": {c: row[c][3] for c in row}, "phases_best": row[best]}), flush=True)
