#!/usr/bin/env python3
"""Device phases of a table-path MSM with and without the window expansion (zk_table_precompute)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
ctx = zk.Context(0); ctx.set_profiling(True)
rng = np.random.default_rng(1)
for logn in [int(a) for a in sys.argv[1:]] or [20, 16]:
    n = 1 << logn
    tab = zk.PointTable(ctx, n).append_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8))
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ref = None
    for c in [None] + ([15, 16, 17, 18, 19, 20] if logn >= 18 else [11, 13, 15, 16, 17]):
        if c is not None:
            t0 = time.perf_counter(); tab.precompute(c); tp = time.perf_counter() - t0
        acc = np.zeros(4)
        for i in range(5):
            r = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
            if i >= 2: acc += np.array(ctx.last_phase_ms())
        acc /= 3
        ref = ref or bytes(r); assert bytes(r) == ref
        print(json.dumps({"logn": logn, "precomp_c": c, "sort": round(acc[1], 3), "accum": round(acc[2], 3), "reduce": round(acc[3], 3),
                          "sum": round(acc[1:].sum(), 3), "build_ms": None if c is None else round(tp * 1e3, 1)}), flush=True)
