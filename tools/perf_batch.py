#!/usr/bin/env python3
"""Throughput of batches of independent MSMs (m MSMs x per terms) over a shared cached table and from compressed points."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
ctx = zk.Context(0); ctx.set_profiling(True)
rng = np.random.default_rng(1)
for m, per in ((1024, 256), (1024, 1024), (1024, 4096), (256, 16384), (64, 65536)):
    n = m * per
    tab = zk.PointTable(ctx, per).append_uniform(rng.integers(0, 256, size=(per, 64), dtype=np.uint8))
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    seg = np.arange(0, n + 1, per, dtype=np.uint64)
    best = 1e9
    for i in range(4):
        t0 = time.perf_counter(); r = zk.batch_vartime_multiscalar_mul(ctx, sc, tab, seg); best = min(best, time.perf_counter() - t0)
        ph = ctx.last_phase_ms()
    t0 = time.perf_counter(); one = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc[:per], tab); t1 = time.perf_counter() - t0
    assert bytes(one) == bytes(r[0])
    print(json.dumps({"m": m, "terms_each": per, "batch_ms": round(best * 1e3, 3), "msm_per_s": round(m / best), "points_per_s": round(n / best),
                      "one_at_a_time_ms_each": round(t1 * 1e3, 3),
                      "device_ms": round(sum(ph[1:]), 3), "device_phases": [round(x, 3) for x in ph[1:]], "device_msm_per_s": round(m / (sum(ph[1:]) * 1e-3))}), flush=True)
