#!/usr/bin/env python3
"""Phase timings (CUDA events inside the library) for table-path and compressed-path MSMs.  usage: perf_phases.py [logn ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
ctx = zk.Context(0); ctx.set_profiling(True)
rng = np.random.default_rng(1)
for logn in [int(a) for a in sys.argv[1:]] or [20]:
    n = 1 << logn
    tab = zk.PointTable(ctx, n).append_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8))
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    acc = np.zeros(4); reps = 6
    for i in range(reps + 2):
        zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
        if i >= 2: acc += np.array(ctx.last_phase_ms())
    acc /= reps
    print(json.dumps({"lib": os.path.basename(os.environ.get("ZKMSM_LIB", "libzkmsm.so")), "logn": logn, "c": zk.pick_window(n),
                      "sort_ms": round(acc[1], 4), "accum_ms": round(acc[2], 4), "reduce_ms": round(acc[3], 4), "sum_ms": round(acc[1:].sum(), 4)}))
