#!/usr/bin/env python3
"""Batch of m MSMs over one shared table, with and without the window expansion, for several widths (device phases)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
ctx = zk.Context(0); ctx.set_profiling(True)
rng = np.random.default_rng(1)
for m, per in ((1024, 4096), (1024, 1024), (256, 16384)):
    n = m * per
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    seg = np.arange(0, n + 1, per, dtype=np.uint64)
    tab = zk.PointTable(ctx, per).append_uniform(rng.integers(0, 256, size=(per, 64), dtype=np.uint8))
    ref = None
    for c in [None, 8, 9, 10, 11, 12, 13, 14, 15]:
        if c is not None: tab.precompute(c)
        for i in range(3): r = zk.batch_vartime_multiscalar_mul(ctx, sc, tab, seg)
        ph = ctx.last_phase_ms()
        h = b"".join(bytes(x) for x in r); ref = ref or h; assert h == ref
        print(json.dumps({"m": m, "terms": per, "precomp_c": c, "device_ms": round(sum(ph[1:]), 3), "phases": [round(x, 3) for x in ph[1:]],
                          "msm_per_s": round(m / (sum(ph[1:]) * 1e-3))}), flush=True)
