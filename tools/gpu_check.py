#!/usr/bin/env python3
"""Developer GPU check (not a test, not the bench): parity spot checks against the pure-Python oracle,
phase timings and the integer-pipe microbenchmarks.  Run under gpurun; writes gpurun_out/gpu_check.log."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zkvm_b200 as zk
from oracle import ristretto255_ref as ref

out = open("gpurun_out/gpu_check.log", "a") if os.path.isdir("gpurun_out") else sys.stdout
def log(**kw):
    s = json.dumps(kw); print(s, flush=True)
    if out is not sys.stdout: out.write(s + "\n"); out.flush()

ctx = zk.Context(0)
for kind, name in ((0, "imad_wide_per_s"), (1, "imad32_per_s"), (2, "fe_mul_per_s"), (3, "fe_sqr_per_s")):
    log(bench=name, value=ctx.bench_int_pipe(kind))

rng = np.random.default_rng(7)
# hash-to-group + encode + decode round trip vs oracle
u = rng.integers(0, 256, size=(64, 64), dtype=np.uint8)
t = zk.PointTable(ctx, 64).append_uniform(u)
enc = t.compress()
want = b"".join(ref.from_uniform_bytes(bytes(r)).encode() for r in u)
log(check="from_uniform+encode", ok=enc == want)
t2 = zk.PointTable(ctx).append_compressed(enc)
log(check="decode->encode roundtrip", ok=t2.compress() == enc)
bad = bytearray(enc); bad[32 * 5] ^= 1
try:
    zk.PointTable(ctx).append_compressed(bytes(bad)); log(check="reject negative s", ok=False)
except zk.InvalidPoint as e:
    log(check="reject negative s", ok=e.index == 5)

pts = [enc[32 * i:32 * i + 32] for i in range(64)]
for n in (0, 1, 2, 17, 64):
    sc = [bytes(rng.integers(0, 256, size=32, dtype=np.uint8)) for _ in range(n)]
    want = ref.msm_naive(sc, pts[:n])
    for c in (0, 4, 7, 8, 11, 13, 16):
        ctx.set_window(c)
        got = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts[:n])
        ok = got is not None and bytes(got) == want
        if not ok or c == 0: log(check=f"msm n={n} c={c}", ok=ok)
ctx.set_window(0)

# large: same answer for every window width; phase timings
ctx.set_profiling(True)
for logn in (10, 14, 16, 18, 20):
    n = 1 << logn
    u = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    tab = zk.PointTable(ctx, n).append_uniform(u)
    comp = tab.compress()
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    res = {}
    for c in ((8, 10, 12, 13, 14, 15, 16) if logn >= 14 else (5, 6, 7, 8, 9, 10, 12)):
        ctx.set_window(c)
        for rep in range(2):
            t0 = time.perf_counter(); r = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab); dt = time.perf_counter() - t0
        res[c] = bytes(r).hex()
        log(msm="table", logn=logn, c=c, wall_ms=dt * 1e3, phases_ms=ctx.last_phase_ms())
    log(check=f"all windows agree logn={logn}", ok=len(set(res.values())) == 1, value=list(res.values())[0])
    ctx.set_window(0)
    for rep in range(2):
        t0 = time.perf_counter(); r2 = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, comp); dt = time.perf_counter() - t0
    log(msm="compressed", logn=logn, c=zk.pick_window(n), wall_ms=dt * 1e3, phases_ms=ctx.last_phase_ms(),
        ok=bytes(r2).hex() == list(res.values())[0])
log(done=True, launches=ctx.launch_count)
