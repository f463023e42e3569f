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
`e2e`    : the same metric through the C ABI with HOST buffers (pinned): per step 32 B scalar + 32 B compressed point
           per term go host->device, are decoded on the device, and the 32-byte result comes back -- all inside the
           timed region.  N = 1: zk_msm_vartime().  N > 1: one process per GPU calls zk_msm_vartime() on its shard, the
           N 32-byte partial encodings are gathered and rank 0 adds them (on its GPU) into the final encoding;
           `e2e.single_process` is the same workload through ONE call of zk_mgpu_msm_vartime() from rank 0 (what a Rust
           verifier would use), `e2e.pageable` the N = 1 call from ordinary pageable memory.
Every timed path is first compared with the CPU oracle ON THE SAME BYTES (`parity`); the sweep n = 2^10..2^20
(BASELINE config 2) and the fixed-total block-scale sizes 2^22 / 2^23 (config 5, `strong`) are parity-gated too.
The "verified ZkVM tx/s" half of BASELINE.json's metric is blocked (needs the slingshot zkvm/bulletproofs
sources) and is reported as such, not estimated.
"""
import argparse
import json
import os
import queue
import subprocess
import sys
import threading
import time

# Several contexts (4 streams each) are in flight per GPU; with the driver's default of 8 hardware queues their streams
# alias and serialise falsely.  libzkmsm sets the same default when it is loaded; stated here too because the variable
# must be in place before the process creates its CUDA context.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_N = 20
METRIC = "ristretto255_vartime_msm_points_per_s"
UNIT = "points/s"
MAC_PER_FE_MUL = 72      # 64 limb products + 8 for the 2^256 = 38 fold
# dram__bytes_read.sum + dram__bytes_write.sum of k_bucket_accum at n = 2^20, c = 16, from one `ncu --set full` capture
NCU_ACCUM_SUMMARY = "profiles/r02_ncu_full_k_bucket_accum.txt"   # written by tools/summarise_ncu.py from the .ncu-rep


def ncu_accum_dram():
    """DRAM bytes (read + write) of one k_bucket_accum launch, parsed from the committed ncu summary of this build."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, seen = 0.0, 0
    try:
        for line in open(os.path.join(ROOT, NCU_ACCUM_SUMMARY)):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1].replace(",", "")) * unit[f[2]]; seen += 1
    except (OSError, KeyError, ValueError):
        return None
    return {"bytes": tot, "source": NCU_ACCUM_SUMMARY} if seen == 2 else None


BLOCKED = "verified ZkVM tx/s: blocked, needs slingshot zkvm + bulletproofs + dalek sources (SURVEY.md section 0)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2_N, help="points per GPU (default 2^20, the headline config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--window", type=int, default=0, help="tuning only: force the Pippenger window width (0 = the library's rule)")
    ap.add_argument("--no-extras", action="store_true", help="skip sweep / strong / batch / shapes (headline numbers only)")
    ap.add_argument("--e2e-threads", type=int, default=0,
                    help="host threads of the end-to-end loop (0 = one per context in flight, fewer when cores are scarce and spinning)")
    ap.add_argument("--inflight", type=int, default=4,
                    help="MSMs in flight: each uses its own context (stream + workspace), so the serial tail of one step "
                         "(Horner + encode, a few warps) overlaps the next step's accumulation; 1 = strictly serial")
    return ap.parse_args()


def config_for(log2n, gpus):
    """Workload description shared verbatim by both arms (the driver compares them)."""
    return {"workload": f"raw ristretto255 vartime MSM, n=2^{log2n} points per GPU x {gpus} GPU(s), uniform scalars mod l, hash-to-group points",
            "log2_points_per_gpu": log2n, "gpus": gpus,
            "inputs": "32-byte scalars (uniform 256-bit strings, used mod l) and 32-byte compressed ristretto255 points (RFC 9496 hash-to-group of seeded bytes)",
            "l2": "2 alternating resident input sets of 128 MiB each (> 126 MB L2); workspace ~300 MiB per context",
            "parallelism": f"point-range shards x{gpus}, one gather of partial points per step" if gpus > 1 else "single GPU"}


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
    """CPU-only input synthesis (reference arm): uniform 32-byte scalars, points = hash-to-group of seeded bytes."""
    import numpy as np
    from oracle import c_oracle
    rng = np.random.default_rng(seed)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    return sc, pts


def cpu_time_msm(sc, pts, n, threads, vector=False):
    """(seconds for decode, seconds for the MSM over decoded points + encode, result bytes).  vector=True: the port's
    4-lane AVX-512 IFMA bucket accumulation (dalek's vector-backend shape) where the host CPU has it; the default is the
    serial radix-2^51 code, which is also the only code the parity gates use."""
    from oracle import c_oracle
    t0 = time.perf_counter(); ge, bad = c_oracle.decompress(pts, n, threads); t1 = time.perf_counter()
    assert bad is None
    c_oracle.set_vector(vector)
    try:
        t1 = time.perf_counter(); r = c_oracle.msm_decompressed(sc, ge, n, threads); t2 = time.perf_counter()
    finally:
        c_oracle.set_vector(False)
    return t1 - t0, t2 - t1, r


def cpu_vector_backend():
    """Name of the port's vector backend on this host, or None."""
    from oracle import c_oracle
    have = c_oracle.set_vector(True); c_oracle.set_vector(False)
    return "avx512-ifma, 4 lanes = the 4 coordinates of a point (dalek's vector-backend shape)" if have else None


# ------------------------------------------------------------------------------------------------------------
def run_reference(a):
    """CPU arm: the reference algorithm's port on all host threads.  Rank 0 only under torchrun.
    `value` = MSM over already-decompressed points (what the CUDA arm's `value` measures); `e2e` = decode + MSM + encode
    from compressed points (what the CUDA arm's `e2e` measures).  Each step is a bounded sample of the workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    threads = host_threads()
    n_total = (1 << a.log2n) * a.gpus
    n_step = min(n_total, 1 << 20)           # bounded sample per step: ~0.3 s MSM + ~0.3 s decode on 16 threads
    sc, pts = synth_inputs_cpu(n_step, 2020)
    W, K = max(0, a.warmup), max(1, a.steps)
    vec = cpu_vector_backend()
    for _ in range(W):
        cpu_time_msm(sc, pts, n_step, threads, vector=bool(vec))
    t_dec = t_msm = t_serial = 0.0
    KS = min(K, 3)                               # steps on which the serial backend is timed as well
    for k in range(K):
        d, m, r_fast = cpu_time_msm(sc, pts, n_step, threads, vector=bool(vec))
        t_dec += d; t_msm += m
        if vec and k < KS:                       # the serial backend on the same step, for the record (and as a cross-check)
            _d, ms, r_ser = cpu_time_msm(sc, pts, n_step, threads)
            t_serial += ms
            assert r_ser == r_fast, "vector and serial CPU paths disagree"
    v = n_step * K / t_msm                       # the faster faithful build of the reference algorithm is the arm's value
    e2e = n_step * K / (t_dec + t_msm)
    sample = (f"{n_step} of {n_total} points per step, {K} steps after {W} warm-up, {threads} threads (index-range sharded Straus/Pippenger); "
              "value = MSM over decompressed points + encode, e2e = decode + MSM + encode"
              + ("; bucket accumulation on the vector backend, everything else (decode, bucket reduction, Horner) serial" if vec else ""))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
        "ms_per_step": t_msm / K * 1e3 * (n_total / n_step), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (radix-2^51 limbs)", "data": "synthetic",
        "config": config_for(a.log2n, a.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "e2e_points_per_s": e2e, "vector_backend": vec,
                         "serial_backend_points_per_s": (n_step * KS / t_serial) if vec else v,
                         "note": "C restatement of dalek's Straus/Pippenger (oracle/msm_oracle.c), serial radix-2^51 backend and, where the "
                                 "host has AVX-512 IFMA, a 4-lane vector backend for the bucket accumulation; dalek itself is not buildable "
                                 "here (no source, no Rust)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "blocked": BLOCKED}))


# ------------------------------------------------------------------------------------------------------------
def run_cuda(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import zkvm_b200 as zk
    from oracle import c_oracle                      # the CHECKER: parity gates before timing, and the cpu_baseline leg
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
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_pg = dist.new_group(backend="gloo")      # host-side barriers and 32-byte gathers: no GPU work while waiting
    c_oracle.load()
    cpu_threads = max(1, host_threads() // world)
    F = max(1, a.inflight)
    # host threads of the end-to-end loop: one per context in flight, but never more than the cores this rank can have
    # (the driver's 8-GPU box exposes 32 cores: 8 ranks x 4 callers + 8 main threads would oversubscribe it)
    cores_per_rank = host_threads() // world
    scarce_cores = cores_per_rank < F + 4            # few cores per GPU: the e2e loop's waiting callers must not spin
    FE = a.e2e_threads if a.e2e_threads > 0 else F   # blocked callers cost no CPU, so scarce cores do not limit their number
    ctxs = [zk.Context(local) for _ in range(max(F, FE))]       # `value` uses the first F, the e2e loop the first FE
    ctx = ctxs[0]
    if a.window:
        for cx in ctxs: cx.set_window(a.window)
    streams = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]
    gather_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in ctxs] if world > 1 else []
    n = 1 << a.log2n
    K, W = a.steps, max(a.warmup, 3)

    def host_barrier():
        if world > 1: dist.barrier(group=host_pg)

    def barrier():
        torch.cuda.synchronize()
        host_barrier()
        torch.cuda.synchronize()

    def gen_inputs(count, seed):
        """Host bytes first (they are what the oracle is handed); points = GPU hash-to-group of seeded bytes, encoded."""
        rng = np.random.default_rng(seed)
        u = rng.integers(0, 256, size=(count, 64), dtype=np.uint8)
        t = zk.PointTable(ctx, count).append_uniform(u)
        comp = np.frombuffer(t.compress(), dtype=np.uint8).copy()
        sc = rng.integers(0, 256, size=(count, 32), dtype=np.uint8).reshape(-1)
        return sc, comp, t

    # ---- synthetic inputs: generated on the host, the SAME bytes go to the GPU arm and to the oracle ----
    SETS = 2   # alternate between two resident input sets so consecutive steps share nothing in L2
    tables, scal_dev, comp_host, scal_host = [], [], [], []
    for s in range(SETS):
        sc, comp, t = gen_inputs(n, 2020 + 16 * rank + s)
        tables.append(t)
        scal_host.append(torch.from_numpy(sc).pin_memory()); comp_host.append(torch.from_numpy(comp).pin_memory())
        scal_dev.append(scal_host[s].to(dev))
    torch.cuda.synchronize()
    np_scal = [t.numpy() for t in scal_host]; np_comp = [t.numpy() for t in comp_host]

    parts = [torch.empty(PARTIAL_BYTES, dtype=torch.uint8, device=dev) for _ in range(F)]
    gathered = [torch.empty(world, PARTIAL_BYTES, dtype=torch.uint8, device=dev) for _ in range(F)]

    def make_runner(tabs, scals, count):
        def submit(i):
            """Queue step i on context i % F: the whole MSM pipeline, asynchronously, plus (N > 1) the one gather."""
            f, s = i % F, i % len(tabs)
            ctxs[f].msm_table_dev(scals[s].data_ptr(), tabs[s], 0, count, parts[f].data_ptr())
            if world > 1:
                # the 128-byte gather is latency-critical and tiny: issue it at high priority, ordered after the MSM and
                # before whatever the context's stream does next, so it is not queued behind other contexts' bulk grids
                gather_streams[f].wait_stream(streams[f])
                with torch.cuda.stream(gather_streams[f]):
                    dist.all_gather_into_tensor(gathered[f], parts[f].view(1, PARTIAL_BYTES))
                streams[f].wait_stream(gather_streams[f])

        def collect(i):
            """Finish step i: sum the partial(s), encode, read the 32 bytes back (synchronises that context only)."""
            f = i % F
            if world > 1:
                return ctxs[f].ext_sum_compress_dev(gathered[f].data_ptr(), world) if rank == 0 else ctxs[f].sync()
            return ctxs[f].ext_sum_compress_dev(parts[f].data_ptr(), 1)

        def run_steps(k):
            out = []
            for i in range(k):
                submit(i)
                if i >= F - 1: out.append(collect(i - (F - 1)))
            for i in range(max(0, k - (F - 1)), k): out.append(collect(i))
            return out
        return run_steps

    run_steps = make_runner(tables, scal_dev, n)

    def timed(fn, k):
        """CUDA events on torch's current stream bracketing work issued on the contexts' streams; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record(cur)
        for st in streams: st.wait_event(e0)
        out = fn(k)
        for st in streams: cur.wait_stream(st)
        e1.record(cur)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    # ---- parity gates before timing: the oracle on the same bytes --------------------------------------------------
    def oracle_total(sc_np, comp_np, count):
        """Oracle encoding of the WHOLE job: every rank's oracle partial, gathered on the host, summed by the oracle."""
        mine = c_oracle.msm(sc_np, comp_np, count, threads=cpu_threads)
        assert mine is not None
        if world == 1:
            return mine
        lst = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(lst, torch.frombuffer(bytearray(mine), dtype=torch.uint8), group=host_pg)
        return c_oracle.point_sum(b"".join(bytes(t.tolist()) for t in lst), world)

    t0 = time.perf_counter()
    want = [oracle_total(np_scal[s], np_comp[s], n) for s in range(SETS)]
    oracle_s = time.perf_counter() - t0
    parity = {"checked_against": "oracle (oracle/msm_oracle.c) on identical host bytes", "n": n * world, "world": world, "paths": {}}

    def gate(name, got, expect):
        ok = (rank != 0) or (got is not None and bytes(got) == bytes(expect))
        parity["paths"][name] = bool(ok)
        if not ok:
            raise SystemExit(f"parity gate failed before timing: {name} disagrees with the oracle")

    r = run_steps(SETS)
    for s in range(SETS): gate(f"table_set{s}", r[s] if rank == 0 else None, want[s])

    # the host-buffer path (per process) with the combine actually finished on rank 0
    comb_ctx = None
    if world > 1 and rank == 0:
        comb_ctx = zk.Context(local); comb_ctx.set_priority(True)     # its one small kernel must not queue behind bulk work

    def finish_gather(enc):
        """N > 1: gather the N 32-byte partial encodings on the host; rank 0 adds them on its GPU (zk_sum_compressed:
        decode + add + encode in one small launch)."""
        if world == 1: return enc
        lst = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(lst, torch.frombuffer(bytearray(bytes(enc)), dtype=torch.uint8), group=host_pg)
        if rank != 0: return None
        return comb_ctx.sum_compressed(np.concatenate([t.numpy() for t in lst]))

    for s in range(SETS):
        gate(f"host_compressed_set{s}", finish_gather(zk.RistrettoPoint.optional_multiscalar_mul(ctx, np_scal[s], np_comp[s])), want[s])

    # ---- `value`: inputs resident in HBM, CUDA events on the launching streams ----
    run_steps(W)
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    launches0 = sum(c.launch_count for c in ctxs[:F])
    ms, outs = timed(run_steps, K)
    launches = sum(c.launch_count for c in ctxs[:F]) - launches0
    if rank == 0:
        for i, o in enumerate(outs): gate("timed_steps", o, want[i % SETS])
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
        tpc = time.perf_counter()
        for tb in tables: tb.precompute(0)
        ctx.sync(); build_ms = (time.perf_counter() - tpc) * 1e3 / SETS
        r = run_steps(max(W, SETS))
        for s in range(SETS): gate(f"precomputed_set{s}", r[s] if rank == 0 else None, want[s])
        pre_ms, _ = timed(run_steps, K)
        precomp = {"what": "same steps over window-expanded static tables (per-window multiples cached in HBM, no doublings, shared buckets)",
                   "value": n * world * K / (pre_ms * 1e-3), "ms_per_step": pre_ms / K, "window_bits": tables[0].precomputed_window,
                   "table_bytes_per_point": 96 * ((254 + tables[0].precomputed_window - 1) // tables[0].precomputed_window),
                   "build_ms_per_table": build_ms}
        for s in range(SETS):                 # drop the expansions: the sections below use the plain tables
            tables[s].clear(); tables[s].append_compressed(np_comp[s])

    # ---- roofline of the dominant kernel (bucket accumulation), measured live with CUDA events ----
    ctx.set_profiling(True)
    acc_ms, phases = [], None
    for i in range(4):
        zk.RistrettoPoint.vartime_multiscalar_mul(ctx, np_scal[i % SETS], tables[i % SETS])
        phases = ctx.last_phase_ms()
        if i >= 1: acc_ms.append(phases[2])
    ctx.set_profiling(False)
    acc_ms = sum(acc_ms) / len(acc_ms)
    c = a.window or zk.pick_window(n); Wn = (254 + c - 1) // c
    adds = n * Wn * (1.0 - 2.0 ** -c)                       # entries = nonzero digits
    nbuckets = Wn * (1 << (c - 1))                          # ~every bucket is non-empty at this density: one task each,
    # whose first entry costs 1 multiply (ge_from_niels), every other entry a 7-multiply mixed add
    macs = ((adds - nbuckets) * 7 + nbuckets * 1) * MAC_PER_FE_MUL
    imad_peak = ctx.bench_int_pipe(0)
    alg_bytes = adds * (96 + 4) + Wn * (1 << (c - 1)) * (128 + 8)

    # ---- `e2e`: host buffers through the public API; F host threads, one context each (ctypes drops the GIL) ----
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=FE)

    def e2e_loop(call, k, finish=finish_gather):
        """call(f, i) -> 32-byte encoding of step i, issued from host thread f; results are finished in step order."""
        qs = [queue.Queue() for _ in range(FE)]
        futs = [pool.submit(lambda f=f: [qs[f].put(call(f, i)) for i in range(f, k, FE)]) for f in range(FE)]
        out = []
        for i in range(k):
            out.append(finish(qs[i % FE].get()))
        for x in futs: x.result()
        return out

    def time_e2e(call, k, finish=finish_gather, sync_ranks=True):
        e2e_loop(call, W, finish)
        if sync_ranks: barrier()
        t0 = time.perf_counter()
        res = e2e_loop(call, k, finish)
        if sync_ranks: barrier()
        dt = time.perf_counter() - t0
        if sync_ranks and world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = t.item()
        return dt, res

    call_pinned = lambda f, i: zk.RistrettoPoint.optional_multiscalar_mul(ctxs[f], np_scal[i % SETS], np_comp[i % SETS])
    if scarce_cores:
        for cx in ctxs: cx.set_wait(1)
    e2e_s, res = time_e2e(call_pinned, K)
    if rank == 0:
        for i, o in enumerate(res): gate("e2e_timed_steps", o, want[i % SETS])
    barrier(); t1 = time.perf_counter()
    for i in range(3): finish_gather(call_pinned(0, i))
    e2e_latency_ms = (time.perf_counter() - t1) / 3 * 1e3
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed regions

    # pageable host memory (what a Rust Vec<u8> is): the library stages it through its pinned ring
    e2e_pageable = None
    if world == 1:
        pg_scal = [np.array(x) for x in np_scal]; pg_comp = [np.array(x) for x in np_comp]
        call_pg = lambda f, i: zk.RistrettoPoint.optional_multiscalar_mul(ctxs[f], pg_scal[i % SETS], pg_comp[i % SETS])
        kp = max(6, K // 2)
        b0 = sum(cx.staged_bytes for cx in ctxs)
        pg_s, res = time_e2e(call_pg, kp)
        staged = sum(cx.staged_bytes for cx in ctxs) - b0
        for i, o in enumerate(res): gate("e2e_pageable_steps", o, want[i % SETS])
        t1 = time.perf_counter()
        for i in range(3): call_pg(0, i)
        e2e_pageable = {"value": n * kp / pg_s, "unit": UNIT, "ms_per_step": pg_s / kp * 1e3, "steps": kp,
                        "single_call_latency_ms": (time.perf_counter() - t1) / 3 * 1e3,
                        "staged_bytes_per_step": staged / (kp + W),
                        "api": "zk_msm_vartime() from pageable numpy arrays: uploads go through the ctx's 4 x 4 MiB pinned staging ring"}
        del pg_scal, pg_comp

    # one process, all N GPUs, one call (zk_mgpu_msm_vartime): what a single-process caller gets
    e2e_single = None
    if world > 1:
        host_barrier()
        try:
            if rank == 0:
                # the whole job's host bytes in ONE buffer on rank 0 (rank r's shard = rank r's seeded inputs)
                big_s = torch.empty(n * world * 32, dtype=torch.uint8).pin_memory(); big_p = torch.empty(n * world * 32, dtype=torch.uint8).pin_memory()
            for r_ in range(world):
                # every rank regenerates nothing: shards travel over the host process group (one-off, outside any timing)
                buf_s = torch.from_numpy(np_scal[0]).clone() if rank == r_ else torch.empty(n * 32, dtype=torch.uint8)
                buf_p = torch.from_numpy(np_comp[0]).clone() if rank == r_ else torch.empty(n * 32, dtype=torch.uint8)
                dist.broadcast(buf_s, src=r_, group=host_pg); dist.broadcast(buf_p, src=r_, group=host_pg)
                if rank == 0:
                    big_s[r_ * n * 32:(r_ + 1) * n * 32] = buf_s; big_p[r_ * n * 32:(r_ + 1) * n * 32] = buf_p
            if rank == 0:
                mgs = [zk.MultiGpu(g=world) for _ in range(FE)]
                if scarce_cores:
                    for m in mgs: m.set_wait(1)
                bs, bp = big_s.numpy(), big_p.numpy()
                call_mg = lambda f, i: mgs[f].optional_multiscalar_mul(bs, bp)
                ks = max(6, K // 2)
                got = call_mg(0, 0)
                gate("single_process_mgpu", got, want[0])
                ident = lambda x: x
                dt, res = time_e2e(call_mg, ks, finish=ident, sync_ranks=False)
                for o in res: gate("single_process_mgpu_steps", o, want[0])
                t1 = time.perf_counter()
                for i in range(3): call_mg(0, i)
                lat = (time.perf_counter() - t1) / 3 * 1e3
                e2e_single = {"value": n * world * ks / dt, "unit": UNIT, "ms_per_step": dt / ks * 1e3, "steps": ks,
                              "single_call_latency_ms": lat, "host_threads": FE, "gather": "peer copies (cudaMemcpyPeerAsync, 128 B per device)",
                              "api": f"zk_mgpu_msm_vartime(mg, scalars_host, points_host, {n * world}, out32): ONE process, {world} GPUs, pinned host buffers"}
                # gather by one ncclAllGather instead of peer copies: ONE handle, calls strictly one after the other
                # (collectives of several communicators over the same GPUs must not be issued concurrently)
                try:
                    t1 = time.perf_counter()
                    for i in range(4): call_mg(0, i)
                    peer_serial = (time.perf_counter() - t1) / 4 * 1e3
                    mgs[0].set_gather("nccl")
                    gate("single_process_mgpu_nccl", call_mg(0, 0), want[0])
                    t1 = time.perf_counter()
                    for i in range(4): call_mg(0, i)
                    e2e_single["gather_serial_call_ms"] = {"peer_copies": peer_serial, "nccl_allgather": (time.perf_counter() - t1) / 4 * 1e3}
                except zk.ZkError as e:
                    e2e_single["nccl_gather"] = f"unavailable: {e}"
                for m in mgs: m.close()
                del big_s, big_p
        except Exception as e:                  # an extra, never a reason to lose the headline line (or to hang the other ranks)
            e2e_single = {"error": repr(e)}
        host_barrier()

    for cx in ctxs: cx.set_wait(0)
    extras = not a.no_extras
    # ---- BASELINE config 2: sweep n = 2^10 .. 2^20 on one GPU, GPU == CPU gated per n ----------------------------
    sweep = None
    if extras and world == 1:
        sweep = []
        cpu_vec = cpu_vector_backend()
        for lg in range(10, min(a.log2n, 20) + 1):
            m = 1 << lg
            sc_m, pc_m = np_scal[0][: 32 * m], np_comp[0][: 32 * m]
            th = min(cpu_threads, max(1, m // 256))
            d1 = m1 = None
            if lg <= 18:
                d1, m1, _ = cpu_time_msm(sc_m, pc_m, m, 1)
            dN, mN, wantm = cpu_time_msm(sc_m, pc_m, m, th)
            if lg <= 14:                                   # short runs: best of 3
                for _ in range(2):
                    x = cpu_time_msm(sc_m, pc_m, m, th); dN, mN = min(dN, x[0]), min(mN, x[1])
            mV = cpu_time_msm(sc_m, pc_m, m, th, vector=True)[1] if cpu_vec else None
            got_t = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc_m, tables[0], n=m)
            got_c = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc_m, pc_m)
            ok = bytes(got_t) == wantm and got_c is not None and bytes(got_c) == wantm
            if not ok: raise SystemExit(f"parity gate failed in the sweep at n = 2^{lg}")
            reps = 20 if lg <= 16 else 8
            gms = 1e9
            for _ in range(reps + 1):                      # best of reps (first one is a warm-up), CUDA events on the ctx stream
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(streams[0])
                ctx.msm_table_dev(scal_dev[0].data_ptr(), tables[0], 0, m, parts[0].data_ptr())
                ctx.ext_sum_compress_dev(parts[0].data_ptr(), 1)
                s1.record(streams[0]); torch.cuda.synchronize()
                if _ > 0: gms = min(gms, s0.elapsed_time(s1))
            ems = 1e9
            for _ in range(reps):
                t1 = time.perf_counter(); zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc_m, pc_m)
                ems = min(ems, (time.perf_counter() - t1) * 1e3)
            sweep.append({"log2n": lg, "window_bits": zk.pick_window(m), "gpu_ms": gms, "gpu_points_per_s": m / (gms * 1e-3),
                          "gpu_e2e_ms": ems, "gpu_e2e_points_per_s": m / (ems * 1e-3),
                          "cpu_1t_points_per_s": (m / m1) if m1 else None, "cpu_1t_e2e_points_per_s": (m / (d1 + m1)) if m1 else None,
                          "cpu_Nt_vector_points_per_s": (m / mV) if mV else None,
                          "cpu_Nt_points_per_s": m / mN, "cpu_Nt_e2e_points_per_s": m / (dN + mN), "cpu_threads": th, "parity_ok": ok})

    # ---- BASELINE config 5: ONE block-scale MSM of fixed total size across the N GPUs (strong scaling) -----------
    strong = None
    if extras and a.log2n >= 20:
        strong = {"what": "one MSM of n_total points sharded by point range over this run's N GPUs (n_total / N per GPU), cached points, "
                          "3 in flight, CUDA events, max over ranks; compare ms_per_step across the N = 1/2/4/8 runs", "sizes": []}
        for lgt in (22, 23):
            cnt = (1 << lgt) // world
            sc_b, comp_b, tab_b = gen_inputs(cnt, 7000 + 16 * rank + lgt)
            sdev = torch.from_numpy(sc_b).to(dev)
            wb = oracle_total(sc_b, comp_b, cnt)
            runb = make_runner([tab_b], [sdev], cnt)
            r = runb(max(3, F))
            ok = rank != 0 or bytes(r[0]) == wb
            if not ok: raise SystemExit(f"parity gate failed at the block-scale size 2^{lgt}")
            ks = 12
            bms, _ = timed(runb, ks)
            strong["sizes"].append({"log2_n_total": lgt, "points_per_gpu": cnt, "ms_per_step": bms / ks,
                                    "points_per_s": (1 << lgt) * ks / (bms * 1e-3), "parity_ok": bool(ok)})
            tab_b.close(); del sdev

    # ---- extra: a batch of independent MSMs (one verdict each) over one cached generator table -----------------
    # Shape stand-in for "verify 1024 transactions": 1024 MSMs x 4096 terms.  NOT a tx/s figure (tx sizes unknown).
    batch = None
    if extras and rank == 0 and world == 1 and a.log2n >= 12:
        bm, bper = 1024, 4096
        bs = torch.randint(0, 256, (bm * bper, 32), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).pin_memory()
        seg = np.arange(0, bm * bper + 1, bper, dtype=np.uint64)
        ctx.set_profiling(True)
        bt = 1e9
        for i in range(3):
            t0 = time.perf_counter(); rb = zk.batch_vartime_multiscalar_mul(ctx, bs.numpy(), tables[0], seg); bt = min(bt, time.perf_counter() - t0)
        ph = ctx.last_phase_ms()
        for k in (0, 511, 1023):                     # spot-check three of the 1024 verdicts against the oracle
            assert bytes(rb[k]) == c_oracle.msm(bs.numpy()[k * bper:(k + 1) * bper], np_comp[0][: 32 * bper], bper, threads=cpu_threads)
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
    if extras and rank == 0 and world == 1 and a.log2n >= 16:
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

    e2e_ms = e2e_s * 1e3
    if rank == 0:
        peaks = {}
        try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception: pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        NCU_ACCUM_DRAM = ncu_accum_dram() if (a.log2n == 20 and c == 16) else None
        traffic = NCU_ACCUM_DRAM["bytes"] if NCU_ACCUM_DRAM else None
        parity["ok"] = all(parity["paths"].values())
        parity["oracle_seconds"] = oracle_s
        e2e = {"value": n * world * K / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / K,
               "h2d_bytes_per_step": 64 * n * world, "d2h_bytes_per_step": 32 * world, "single_call_latency_ms": e2e_latency_ms,
               "host_threads": FE, "host_cores_per_rank": cores_per_rank, "wait": "blocking" if scarce_cores else "spin",
               "api": "zk_msm_vartime(ctx, scalars_host, compressed_points_host, n, out32) from pinned host memory"
                      + ("; one process per GPU, partial encodings gathered on the host, rank 0 adds them on its GPU" if world > 1 else "")}
        if e2e_pageable: e2e["pageable"] = e2e_pageable
        if e2e_single: e2e["single_process"] = e2e_single
        out = {
            "metric": METRIC, "value": n * world * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8 saturated 32-bit limbs, IMAD.WIDE)", "data": "synthetic",
            "config": config_for(a.log2n, world),
            "impl_config": {"window_bits": c, "windows": Wn, "inflight": F, "single_msm_latency_ms": latency_ms,
                            "pipelining": f"{F} contexts (stream + workspace each) in flight; results are collected in order, one step behind",
                            "stream_priorities": "per context: digit sort, tree, Horner and Encode on a high-priority stream, accumulate/decode at low priority",
                            "cuda_device_max_connections": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"),
                            "value_inputs": "scalars + cached decompressed points (affine Niels, 96 B) resident in HBM"},
            "parity": parity,
            "e2e": e2e,
            "gpu_launches": launches * world,
            "clocks": clocks,
            "roofline": {"bound": "imad", "kernel": "k_bucket_accum", "achieved": macs / (acc_ms * 1e-3) / 1e12, "peak": imad_peak / 1e12,
                         "unit": "T(32x32+64 MAC)/s", "frac": macs / (acc_ms * 1e-3) / imad_peak,
                         "step_frac": macs / (ms / K * 1e-3) / imad_peak,
                         "traffic": traffic, "traffic_source": NCU_ACCUM_DRAM["source"] if traffic else None,
                         "peak_source": "measured live: zk_bench_int_pipe(0), IMAD.WIDE.U32 carry chains on all SMs",
                         "kernel_ms": acc_ms, "phases_ms": {"decompress": phases[0], "digits_sort": phases[1], "bucket_accum": phases[2],
                                                            "reduce_encode": phases[3]},
                         "algorithmic": f"({adds:.0f} entries - {nbuckets} task starts) x 7 fe_mul + {nbuckets} x 1 fe_mul, x {MAC_PER_FE_MUL} MAC each; "
                                        "step_frac = the same MACs over the whole timed step"},
            "roofline_hbm": {"bound": "hbm", "kernel": "k_bucket_accum", "achieved": alg_bytes / (acc_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": alg_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak,
                             "traffic": traffic, "traffic_source": NCU_ACCUM_DRAM["source"] if traffic else None,
                             "algorithmic_bytes": alg_bytes, "peak_source": hbm_src},
            "blocked": BLOCKED,
        }
        if batch: out["batch"] = batch
        if precomp: out["precomputed_tables"] = precomp
        if shapes: out["proof_sized_shapes"] = shapes
        if sweep: out["sweep"] = sweep
        if strong: out["strong"] = strong
        if not a.no_cpu_baseline and world == 1:
            try:
                ns = min(n, 1 << 20)
                d1, m1, _ = cpu_time_msm(np_scal[0][: 32 * (ns // 8)], np_comp[0][: 32 * (ns // 8)], ns // 8, 1)
                dN, mN, rN = cpu_time_msm(np_scal[0][: 32 * ns], np_comp[0][: 32 * ns], ns, cpu_threads)
                vec = cpu_vector_backend()
                mV = rV = None
                if vec:
                    _dv, mV, rV = cpu_time_msm(np_scal[0][: 32 * ns], np_comp[0][: 32 * ns], ns, cpu_threads, vector=True)
                best = min(mN, mV) if mV else mN
                out["cpu_baseline"] = {"value": ns / best, "unit": UNIT, "cores": cpu_threads, "kind": "port",
                                       "sample": f"{ns} points (the GPU arm's own input bytes), MSM over decompressed points + encode, {cpu_threads} threads; "
                                                 f"single-thread figures on {ns // 8} points (serial backend)",
                                       "vector_backend": vec, "serial_backend_points_per_s": ns / mN,
                                       "vector_backend_points_per_s": (ns / mV) if mV else None,
                                       "e2e_points_per_s": ns / (dN + best), "single_thread_points_per_s": (ns // 8) / m1,
                                       "single_thread_e2e_points_per_s": (ns // 8) / (d1 + m1),
                                       "same_bytes_as_gpu": ns != n or (rN == want[0] and (rV is None or rV == want[0]))}
            except Exception as e:      # the baseline is reported, never required for the GPU number
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out))
    if world > 1:
        host_barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)
