#!/usr/bin/env python3
"""bench.py -- Ristretto255 vartime MSM throughput (BASELINE.json metric, config[1]: n = 2^20 on 1 B200).

  python bench.py [--gpus N --steps K --warmup W]            the CUDA path (this repo)
  python bench.py --impl reference [...]                     the CPU path of the reference algorithm
      (oracle/msm_oracle.c, a "port": the reference tree has no source to build)

A step = one pass of the hot path over one batch: an n-point MSM (n = 2^20 per GPU) reduced to one 32-byte
encoding.  At N > 1 the path shards by point range (weak scaling: every rank owns 2^20 points of an
N*2^20-point MSM) and the only collective is one all_gather of 128 bytes per rank per step.

`value`  : points/s with scalars and the decompressed point cache resident in HBM (what dalek's
           vartime_multiscalar_mul is handed), timed with CUDA events on the launching stream.
`e2e`    : the same metric through the C ABI call zk_msm_vartime() with pinned HOST buffers: per step 32 B
           scalar + 32 B compressed point per term go host->device, are decoded on the device, and the 32-byte
           result comes back -- all inside the timed region.
The "verified ZkVM tx/s" half of BASELINE.json's metric is blocked (needs the slingshot zkvm/bulletproofs
sources) and is reported as such, not estimated.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_N = 20
METRIC = "ristretto255_vartime_msm_points_per_s"
UNIT = "points/s"
MAC_PER_FE_MUL = 72      # 64 limb products + 8 for the 2^256 = 38 fold
# dram__bytes_read.sum + dram__bytes_write.sum of k_bucket_accum at n = 2^20, c = 16 (profiles/r01_ncu_full_k_bucket_accum.txt)
NCU_ACCUM_DRAM_BYTES = 1.152978e9 + 59.066368e6
BLOCKED = "verified ZkVM tx/s: blocked, needs slingshot zkvm + bulletproofs + dalek sources (SURVEY.md section 0)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_N, help="points per GPU (default 2^20, the headline config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=3,
                    help="MSMs in flight: each uses its own context (stream + workspace), so the serial tail of one step "
                         "(Horner + encode, a few warps) overlaps the next step's accumulation; 1 = strictly serial")
    return ap.parse_args()


def workload_name(log2n, gpus):
    return f"raw ristretto255 vartime MSM, n=2^{log2n} points per GPU x {gpus} GPU(s), uniform scalars mod l, hash-to-group points"


# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try: self.proc.wait(timeout=5)
        except Exception: self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9: continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower() == "active": reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try: return len(os.sched_getaffinity(0))
    except AttributeError: return os.cpu_count() or 1


def synth_inputs_cpu(n, seed):
    """Same distribution as the GPU arm: uniform 32-byte scalars, points = hash-to-group of seeded bytes."""
    import numpy as np
    from oracle import c_oracle
    rng = np.random.default_rng(seed)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    return sc, pts


def cpu_time_msm(sc, pts, n, threads, reps=1):
    """Seconds for decode + MSM + encode (what e2e measures) and for the MSM alone, best of reps."""
    from oracle import c_oracle
    best_full, best_msm = 1e30, 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); ge, bad = c_oracle.decompress(pts, n, threads); t1 = time.perf_counter()
        assert bad is None
        c_oracle.msm_decompressed(sc, ge, n, threads); t2 = time.perf_counter()
        best_full = min(best_full, t2 - t0); best_msm = min(best_msm, t2 - t1)
    return best_full, best_msm


# ------------------------------------------------------------------------------------------------------------
def run_reference(a):
    """CPU arm: the reference algorithm's port on all host threads.  Rank 0 only under torchrun."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    threads = host_threads()
    n_total = (1 << a.log2n) * a.gpus
    n_step = min(n_total, 1 << 21)           # bounded sample per step
    sc, pts = synth_inputs_cpu(n_step, 2020)
    for _ in range(max(1, min(a.warmup, 1))):
        cpu_time_msm(sc[: n_step // 8], pts[: 32 * (n_step // 8)], n_step // 8, threads)
    steps = max(1, min(a.steps, 5))
    t_full = t_msm = 0.0
    for _ in range(steps):
        f, m = cpu_time_msm(sc, pts, n_step, threads)
        t_full += f; t_msm += m
    v = n_step * steps / t_full
    sample = f"{n_step} of {n_total} points per step, {steps} steps, decode+MSM+encode, {threads} threads (index-range sharded Pippenger)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": 1,
        "ms_per_step": t_full / steps * 1e3 * (n_total / n_step), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (radix-2^51 limbs)", "data": "synthetic",
        "config": {"workload": workload_name(a.log2n, a.gpus)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "msm_only_points_per_s": n_step * steps / t_msm,
                         "note": "C restatement of dalek's Straus/Pippenger (oracle/msm_oracle.c); dalek itself is not buildable here (no source, no Rust)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "blocked": BLOCKED}))


# ------------------------------------------------------------------------------------------------------------
def run_cuda(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import zkvm_b200 as zk
    from zkvm_b200.sharded import PARTIAL_BYTES

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {a.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the MSM path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F = max(1, a.inflight)
    ctxs = [zk.Context(local) for _ in range(F)]
    ctx = ctxs[0]
    streams = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]
    n = 1 << a.log2n
    K, W = a.steps, max(a.warmup, 3)

    # ---- synthetic inputs, created on the device for `value`, on pinned host memory for `e2e` ----
    g = torch.Generator(device=dev); g.manual_seed(2020 + rank)
    SETS = 2   # alternate between two resident input sets so consecutive steps share nothing in L2
    tables, scal_dev = [], []
    for s in range(SETS):
        u = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device=dev, generator=g)
        torch.cuda.synchronize()
        t = zk.PointTable(ctx, n).append_uniform_dev(u.data_ptr(), n)
        tables.append(t)
        scal_dev.append(torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g).contiguous())
        del u
    comp_host, scal_host = [], []
    for s in range(SETS):
        c = torch.empty(n * 32, dtype=torch.uint8, device=dev)
        tables[s].compress_dev(c.data_ptr()); ctx.sync()
        comp_host.append(c.cpu().pin_memory()); scal_host.append(scal_dev[s].reshape(-1).cpu().pin_memory())
        del c
    torch.cuda.synchronize()

    parts = [torch.empty(PARTIAL_BYTES, dtype=torch.uint8, device=dev) for _ in range(F)]
    gathered = [torch.empty(world, PARTIAL_BYTES, dtype=torch.uint8, device=dev) for _ in range(F)]

    def submit(i):
        """Queue step i on context i % F: the whole MSM pipeline, asynchronously, plus (N > 1) the one gather."""
        f, s = i % F, i % SETS
        ctxs[f].msm_table_dev(scal_dev[s].data_ptr(), tables[s], 0, n, parts[f].data_ptr())
        if world > 1:
            with torch.cuda.stream(streams[f]):
                dist.all_gather_into_tensor(gathered[f], parts[f].view(1, PARTIAL_BYTES))

    def collect(i):
        """Finish step i: sum the partial(s), encode, read the 32 bytes back (synchronises that context only)."""
        f = i % F
        if world > 1:
            return ctxs[f].ext_sum_compress_dev(gathered[f].data_ptr(), world) if rank == 0 else ctxs[f].sync()
        return ctxs[f].ext_sum_compress_dev(parts[f].data_ptr(), 1)

    def run_steps(k):
        out = None
        for i in range(k):
            submit(i)
            if i >= F - 1: out = collect(i - (F - 1))
        for i in range(max(0, k - (F - 1)), k): out = collect(i)
        return out

    np_scal = [t.numpy() for t in scal_host]; np_comp = [t.numpy() for t in comp_host]

    def step_e2e(i):
        """One public-API call from host buffers on context i % F (blocking; returns the 32-byte encoding)."""
        f, s = i % F, i % SETS
        return zk.RistrettoPoint.optional_multiscalar_mul(ctxs[f], np_scal[s], np_comp[s])

    def gather_e2e(i, r):
        """N > 1: this rank's partial came back as 32 bytes; one gather (issued from the main thread, in step order,
        so every rank enqueues its collectives identically); rank 0 would add the G encodings."""
        f = i % F
        enc = torch.frombuffer(bytearray(bytes(r) + bytes(PARTIAL_BYTES - 32)), dtype=torch.uint8).to(dev)
        dist.all_gather_into_tensor(gathered[f], enc.view(1, PARTIAL_BYTES))

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate before timing: device path == host-buffer path, bit for bit ----
    r_dev = run_steps(1)
    r_e2e = zk.RistrettoPoint.optional_multiscalar_mul(ctx, np_scal[0], np_comp[0])
    if world == 1 and bytes(r_dev) != bytes(r_e2e):
        raise SystemExit("parity gate failed: table path and compressed path disagree")

    # ---- `value`: inputs resident in HBM, CUDA events on the launching streams ----
    run_steps(W)
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    barrier()
    launches0 = sum(c.launch_count for c in ctxs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    for st in streams: st.wait_event(e0)
    run_steps(K)
    for st in streams: cur.wait_stream(st)
    e1.record(cur)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(c.launch_count for c in ctxs) - launches0
    # single-MSM latency (one context, strictly serial) for the record
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(streams[0])
    for i in range(4):
        ctx.msm_table_dev(scal_dev[i % SETS].data_ptr(), tables[i % SETS], 0, n, parts[0].data_ptr())
        ctx.ext_sum_compress_dev(parts[0].data_ptr(), 1)
    l1.record(streams[0])
    barrier()
    latency_ms = l0.elapsed_time(l1) / 4

    # ---- extra: the same loop over window-expanded static tables (zk_table_precompute) --------------------------
    precomp = None
    if a.log2n <= 21:
        for tb in tables: tb.precompute(0)
        run_steps(W)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(cur)
        for st in streams: st.wait_event(p0)
        r_pre = run_steps(K)
        for st in streams: cur.wait_stream(st)
        p1.record(cur)
        barrier()
        pre_ms = p0.elapsed_time(p1)
        if world == 1 and bytes(r_pre) != bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, np_scal[(K - 1) % SETS], np_comp[(K - 1) % SETS])):
            raise SystemExit("parity gate failed: precomputed-table path disagrees")
        precomp = {"what": "same steps over window-expanded static tables (per-window multiples cached in HBM, no doublings, shared buckets)",
                   "value": n * world * K / (pre_ms * 1e-3), "ms_per_step": pre_ms / K, "window_bits": tables[0].precomputed_window,
                   "table_bytes_per_point": 96 * ((254 + tables[0].precomputed_window - 1) // tables[0].precomputed_window)}
        tp = torch.tensor([pre_ms], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        precomp["value"] = n * world * K / (tp.item() * 1e-3); precomp["ms_per_step"] = tp.item() / K
        for tb in tables: tb.clear()          # drop the expansions; rebuild the plain tables for the sections below
        g2 = torch.Generator(device=dev); g2.manual_seed(2020 + rank)
        for s_ in range(SETS):
            u = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device=dev, generator=g2)
            torch.cuda.synchronize()
            tables[s_].append_uniform_dev(u.data_ptr(), n)
            _ = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g2)   # keep the generator in step
            del u

    # ---- roofline of the dominant kernel (bucket accumulation), measured live with CUDA events ----
    ctx.set_profiling(True)
    acc_ms, phases = [], None
    for i in range(4):
        zk.RistrettoPoint.vartime_multiscalar_mul(ctx, np_scal[i % SETS], tables[i % SETS])
        phases = ctx.last_phase_ms()
        if i >= 1: acc_ms.append(phases[2])
    ctx.set_profiling(False)
    acc_ms = sum(acc_ms) / len(acc_ms)
    c = zk.pick_window(n); Wn = (254 + c - 1) // c
    adds = n * Wn * (1.0 - 2.0 ** -c)                       # entries = nonzero digits
    nbuckets = Wn * (1 << (c - 1))                          # ~every bucket is non-empty at this density: one task each,
    # whose first entry costs 1 multiply (ge_from_niels), every other entry a 7-multiply mixed add
    macs = ((adds - nbuckets) * 7 + nbuckets * 1) * MAC_PER_FE_MUL
    imad_peak = ctx.bench_int_pipe(0)
    alg_bytes = adds * (96 + 4) + Wn * (1 << (c - 1)) * (128 + 8)

    # ---- `e2e`: host buffers through the public API; F host threads, one context each (ctypes drops the GIL) ----
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=F)

    import queue

    def run_e2e(k):
        qs = [queue.Queue() for _ in range(F)]
        futs = [pool.submit(lambda f=f: [qs[f].put(step_e2e(i)) for i in range(f, k, F)]) for f in range(F)]
        out = []
        for i in range(k):                      # results in step order
            r = qs[i % F].get()
            if world > 1: gather_e2e(i, r)
            out.append(r)
        for x in futs: x.result()
        if world > 1: torch.cuda.synchronize()
        return out

    run_e2e(W)
    barrier()
    t0 = time.perf_counter()
    res = run_e2e(K)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world == 1 and bytes(res[0]) != bytes(r_dev):
        raise SystemExit("parity gate failed inside the e2e loop")
    barrier(); t1 = time.perf_counter()
    for i in range(3):
        r = step_e2e(i * F)
        if world > 1: gather_e2e(i * F, r); torch.cuda.synchronize()
    e2e_latency_ms = (time.perf_counter() - t1) / 3 * 1e3
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions

    # ---- extra: a batch of independent MSMs (one verdict each) over one cached generator table -----------------
    # Shape stand-in for "verify 1024 transactions": 1024 MSMs x 4096 terms.  NOT a tx/s figure (tx sizes unknown).
    batch = None
    if rank == 0 and a.log2n >= 12:
        bm, bper = 1024, 4096
        bs = torch.randint(0, 256, (bm * bper, 32), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).pin_memory()
        seg = np.arange(0, bm * bper + 1, bper, dtype=np.uint64)
        ctx.set_profiling(True)
        bt = 1e9
        for i in range(3):
            t0 = time.perf_counter(); rb = zk.batch_vartime_multiscalar_mul(ctx, bs.numpy(), tables[0], seg); bt = min(bt, time.perf_counter() - t0)
        ph = ctx.last_phase_ms()
        one = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, bs.numpy()[:bper], tables[0], n=bper)
        assert bytes(one) == bytes(rb[0])
        gtab = zk.PointTable(ctx, bper).append_compressed(np_comp[0][: 32 * bper]).precompute(0)   # the shared generators, window-expanded
        bt2 = 1e9
        for i in range(3):
            t0 = time.perf_counter(); rb2 = zk.batch_vartime_multiscalar_mul(ctx, bs.numpy(), gtab, seg); bt2 = min(bt2, time.perf_counter() - t0)
        ph2 = ctx.last_phase_ms(); ctx.set_profiling(False)
        assert [bytes(x) for x in rb2] == [bytes(x) for x in rb]
        gtab.close()
        batch = {"what": "1024 independent MSMs x 4096 terms over one cached table, one 32-byte result each (zk_msm_vartime_table_batch); "
                         "synthetic stand-in for per-transaction verdicts, not tx/s",
                 "msm_per_s_host_api": bm / bt, "ms_host_api": bt * 1e3, "msm_per_s_device": bm / (sum(ph[1:]) * 1e-3),
                 "device_phases_ms": {"digits_sort": ph[1], "bucket_accum": ph[2], "reduce_encode": ph[3]},
                 "window_expanded_table": {"msm_per_s_host_api": bm / bt2, "msm_per_s_device": bm / (sum(ph2[1:]) * 1e-3)}}

    # ---- extra: single-call latencies at the proof-sized shapes BASELINE.json names (MSM level only) -------------
    shapes = None
    if rank == 0 and a.log2n >= 16:
        def lat(fn, reps=5):
            fn(); t0 = time.perf_counter()
            for _ in range(reps): fn()
            return (time.perf_counter() - t0) / reps * 1e3
        m16 = 1 << 16
        gens = zk.PointTable(ctx, m16).append_compressed(np_comp[0][: 32 * m16])
        s16 = np_scal[0][: 32 * m16]; dyn_s = np_scal[1][: 32 * 64]; dyn_p = np_comp[1][: 32 * 64]
        shapes = {"what": "host-API latency of one MSM, ms; MSM-level stand-ins for the proof shapes in BASELINE.json configs "
                          "(the proofs themselves are blocked, SURVEY.md section 0)",
                  "table_2e16": lat(lambda: zk.RistrettoPoint.vartime_multiscalar_mul(ctx, s16, gens)),
                  "mixed_2e16_static_plus_64_dynamic": lat(lambda: zk.RistrettoPoint.mixed_multiscalar_mul(ctx, s16, gens, dyn_s, dyn_p)),
                  "compressed_4096": lat(lambda: zk.RistrettoPoint.optional_multiscalar_mul(ctx, np_scal[0][: 32 * 4096], np_comp[0][: 32 * 4096]))}
        gens.precompute(0)
        shapes["table_2e16_precomputed"] = lat(lambda: zk.RistrettoPoint.vartime_multiscalar_mul(ctx, s16, gens))
        gens.close()

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()

    if rank == 0:
        peaks = {}
        try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception: pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        out = {
            "metric": METRIC, "value": n * world * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8 saturated 32-bit limbs, IMAD.WIDE)", "data": "synthetic",
            "config": {"workload": workload_name(a.log2n, world), "window_bits": c, "windows": Wn,
                       "l2": "2 alternating resident input sets of 128 MiB each (> 126 MB L2); workspace ~300 MiB per context",
                       "inflight": F, "single_msm_latency_ms": latency_ms,
                       "pipelining": f"{F} contexts (stream + workspace each) in flight; results are collected in order, one step behind",
                       "value_inputs": "scalars + cached decompressed points (affine Niels, 96 B) resident in HBM",
                       "parallelism": f"point-range shards x{world}, one 128 B all_gather per step" if world > 1 else "single GPU"},
            "e2e": {"value": n * world * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
                    "h2d_bytes_per_step": 64 * n * world, "d2h_bytes_per_step": 40 * world, "single_call_latency_ms": e2e_latency_ms,
                    "host_threads": F,
                    "api": "zk_msm_vartime(ctx, scalars_host, compressed_points_host, n, out32) from pinned host memory"},
            "gpu_launches": launches * world,
            "clocks": clocks,
            "roofline": {"bound": "imad", "kernel": "k_bucket_accum", "achieved": macs / (acc_ms * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                         "unit": "T(32x32+64 MAC)/s", "frac": macs / (acc_ms * 1e-3) / imad_peak,
                         "traffic": NCU_ACCUM_DRAM_BYTES if (a.log2n == 20 and c == 16) else None,
                         "peak_source": "measured live: zk_bench_int_pipe(0), IMAD.WIDE.U32 carry chains on all SMs",
                         "kernel_ms": acc_ms, "phases_ms": {"decompress": phases[0], "digits_sort": phases[1], "bucket_accum": phases[2],
                                                            "reduce_encode": phases[3]},
                         "algorithmic": f"({adds:.0f} entries - {nbuckets} task starts) x 7 fe_mul + {nbuckets} x 1 fe_mul, x {MAC_PER_FE_MUL} MAC each"},
            "roofline_hbm": {"bound": "hbm", "kernel": "k_bucket_accum", "achieved": alg_bytes / (acc_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": alg_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": NCU_ACCUM_DRAM_BYTES if (a.log2n == 20 and c == 16) else None, "algorithmic_bytes": alg_bytes, "peak_source": hbm_src},
            "blocked": BLOCKED,
        }
        if batch: out["batch"] = batch
        if precomp: out["precomputed_tables"] = precomp
        if shapes: out["proof_sized_shapes"] = shapes
        if not a.no_cpu_baseline:
            try:
                threads = host_threads()
                ns = min(n, 1 << 20)
                sc, pts = synth_inputs_cpu(ns, 7)
                f1, m1 = cpu_time_msm(sc[: ns // 8], pts[: 32 * (ns // 8)], ns // 8, 1)
                fN, mN = cpu_time_msm(sc, pts, ns, threads)
                out["cpu_baseline"] = {"value": ns / fN, "unit": UNIT, "cores": threads, "kind": "port",
                                       "sample": f"{ns} points, decode+MSM+encode, {threads} threads; single-thread figures on {ns // 8} points",
                                       "msm_only_points_per_s": ns / mN, "single_thread_points_per_s": (ns // 8) / f1,
                                       "single_thread_msm_only_points_per_s": (ns // 8) / m1}
            except Exception as e:      # the baseline is reported, never required for the GPU number
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)
