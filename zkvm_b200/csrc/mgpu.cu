// Single-process multi-GPU layer of the C ABI (include/zkmsm.h, zk_mgpu_*): BASELINE.json config 5, "block-scale batch
// MSM sharded by point range over 2/4/8 B200, partial-point gather", reachable from a host program that cannot launch
// one process per GPU (the Rust verifier).
//
// An MSM is a sum of independent terms, so device r of G takes the contiguous index range shard_range(n, r, G), runs
// the whole single-device pipeline (msm.cu) on its slice with no data-path collective, and leaves one 128-byte extended
// partial in its own HBM.  The G partials are then gathered on device 0 -- either by G-1 peer copies of 128 bytes over
// NVLink (default) or by one ncclAllGather of a communicator made with ncclCommInitAll (zk_mgpu_set_gather(mg, 1);
// libnccl is dlopen()ed, never linked) -- and device 0 adds them and encodes (k_ext_sum_encode).  The exchange is
// G*128 bytes: latency-bound, not bandwidth-bound.
//
// Host side: one worker thread per device queues that device's uploads and kernels (uploads from pageable memory
// are synchronous memcpy()s into the pinned staging ring, so they must not be issued from one thread for all devices).
#include <condition_variable>
#include <dlfcn.h>
#include <functional>
#include <mutex>
#include <new>
#include <string.h>
#include <thread>
#include <vector>

#include "internal.h"

namespace {

// ---- minimal NCCL surface, resolved at run time --------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;             // ncclSuccess == 0
constexpr int kNcclUint8 = 1;         // ncclUint8 / ncclChar family: ncclInt8 = 0, ncclUint8 = 1 (nccl.h, stable since 2.0)
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load() {
        if (h) return true;
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllGather && GroupStart && GroupEnd;
    }
};

// ---- one worker thread per device ------------------------------------------------------------------------------
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = ZK_OK;
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            std::function<int()> j = std::move(job);
            has_job = false;
            lk.unlock();
            int r = j();
            lk.lock();
            rc = r; done = true;
            cv.notify_all();
        }
    }
    void post(std::function<int()> j) {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(j); has_job = true; done = false;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done; });
        return rc;
    }
};

}  // namespace

struct zk_mgpu {
    int g = 0;
    std::vector<int> devices;
    std::vector<zk_ctx*> ctx;
    std::vector<Worker*> workers;
    uint8_t* gather0 = nullptr;                 // device 0: g x 128 B
    std::vector<uint8_t*> gather_all;           // NCCL mode: every device holds the g x 128 B receive buffer
    std::vector<cudaEvent_t> ev_part;           // partial of device r has landed in gather0
    int gather_mode = 0;                        // 0 = peer copies, 1 = NCCL all-gather
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    char err[256] = {0};
};

// A sharded point cache: append k of n_k points is cut into g index ranges, device r keeps its range of every append in
// its own zk_table, in append order.  Global index -> (device, local row) is therefore monotone per device, and any
// global slice [offset, offset+n) maps to ONE contiguous run of local rows on each device.
struct zk_mgpu_table {
    zk_mgpu* mg = nullptr;
    std::vector<zk_table*> shard;
    struct Seg { size_t base, n; };             // append k covers global indices [base, base + n)
    std::vector<Seg> segs;
    size_t len = 0;
};

static void shard_range(size_t n, int r, int g, size_t* lo, size_t* hi) {
    const size_t base = n / g, extra = n % g;
    *lo = (size_t)r * base + ((size_t)r < extra ? (size_t)r : extra);
    *hi = *lo + base + ((size_t)r < extra ? 1 : 0);
}

// run fn(r) on every device's worker; returns the first non-OK status (INVALID_POINT ranks below hard errors)
static int run_all(zk_mgpu* mg, const std::function<int(int)>& fn) {
    for (int r = 0; r < mg->g; r++) mg->workers[r]->post([&fn, r] { return fn(r); });
    int rc = ZK_OK;
    for (int r = 0; r < mg->g; r++) {
        int x = mg->workers[r]->wait();
        if (x != ZK_OK && (rc == ZK_OK || rc == ZK_ERR_INVALID_POINT)) rc = x;
    }
    return rc;
}

static void set_err(zk_mgpu* mg, const char* what, const char* detail) { snprintf(mg->err, sizeof(mg->err), "%s: %s", what, detail ? detail : ""); }

extern "C" int zk_mgpu_create(const int* devices, int g, zk_mgpu** out) {
    if (!out || g < 1 || g > 64) return ZK_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        fprintf(stderr, "zkmsm: zk_mgpu_create: no CUDA device (there is no CPU fallback)\n");
        return ZK_ERR_CUDA;
    }
    zk_mgpu* mg = new (std::nothrow) zk_mgpu();
    if (!mg) return ZK_ERR_NOMEM;
    mg->g = g;
    for (int r = 0; r < g; r++) {
        int d = devices ? devices[r] : r;
        if (d < 0 || d >= ndev) { delete mg; return ZK_ERR_ARG; }
        for (int q = 0; q < r; q++) if (mg->devices[q] == d) { delete mg; return ZK_ERR_ARG; }
        mg->devices.push_back(d);
    }
    int rc = ZK_OK;
    for (int r = 0; r < g && rc == ZK_OK; r++) {
        zk_ctx* c = nullptr;
        rc = zk_ctx_create(mg->devices[r], &c);
        if (rc == ZK_OK) mg->ctx.push_back(c);
    }
    if (rc == ZK_OK) {
        // peer access device 0 <-> device r, where the topology offers it (NVLink/NVSwitch on a B200 box); without it
        // cudaMemcpyPeerAsync still works, staged through the host
        for (int r = 1; r < g; r++) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, mg->devices[r], mg->devices[0]) == cudaSuccess && can) {
                cudaSetDevice(mg->devices[r]);
                cudaError_t e = cudaDeviceEnablePeerAccess(mg->devices[0], 0);
                if (e != cudaSuccess) cudaGetLastError();      // already enabled is fine
            }
        }
        cudaSetDevice(mg->devices[0]);
        if (cudaMalloc((void**)&mg->gather0, (size_t)g * 128) != cudaSuccess) rc = ZK_ERR_NOMEM;
        for (int r = 0; r < g && rc == ZK_OK; r++) {
            cudaEvent_t e;
            cudaSetDevice(mg->devices[r]);
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) rc = ZK_ERR_CUDA; else mg->ev_part.push_back(e);
        }
    }
    if (rc == ZK_OK) {
        for (int r = 0; r < g; r++) {
            Worker* w = new (std::nothrow) Worker();
            if (!w) { rc = ZK_ERR_NOMEM; break; }
            w->th = std::thread([w] { w->loop(); });
            mg->workers.push_back(w);
        }
    }
    if (rc != ZK_OK) { zk_mgpu_destroy(mg); return rc; }
    *out = mg;
    return ZK_OK;
}

extern "C" void zk_mgpu_destroy(zk_mgpu* mg) {
    if (!mg) return;
    for (Worker* w : mg->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (size_t r = 0; r < mg->comms.size(); r++) if (mg->comms[r]) mg->nccl.CommDestroy(mg->comms[r]);
    for (size_t r = 0; r < mg->gather_all.size(); r++) if (mg->gather_all[r]) { cudaSetDevice(mg->devices[r]); cudaFree(mg->gather_all[r]); }
    for (size_t r = 0; r < mg->ev_part.size(); r++) { cudaSetDevice(mg->devices[r]); cudaEventDestroy(mg->ev_part[r]); }
    if (mg->gather0) { cudaSetDevice(mg->devices[0]); cudaFree(mg->gather0); }
    for (zk_ctx* c : mg->ctx) zk_ctx_destroy(c);
    delete mg;
}

extern "C" int zk_mgpu_device_count(const zk_mgpu* mg) { return mg ? mg->g : 0; }
extern "C" const char* zk_mgpu_last_error(const zk_mgpu* mg) {
    if (!mg) return "";
    if (mg->err[0]) return mg->err;
    for (zk_ctx* c : mg->ctx) if (c->err[0]) return c->err;
    return "";
}
extern "C" int zk_mgpu_set_staging(zk_mgpu* mg, int mode) {
    if (!mg) return ZK_ERR_ARG;
    for (zk_ctx* c : mg->ctx) TRY(zk_ctx_set_staging(c, mode));
    return ZK_OK;
}
extern "C" int zk_mgpu_set_wait(zk_mgpu* mg, int mode) {
    if (!mg) return ZK_ERR_ARG;
    for (zk_ctx* c : mg->ctx) TRY(zk_ctx_set_wait(c, mode));
    return ZK_OK;
}
extern "C" uint64_t zk_mgpu_launch_count(const zk_mgpu* mg) {
    uint64_t s = 0;
    if (mg) for (zk_ctx* c : mg->ctx) s += c->launches;
    return s;
}

extern "C" int zk_mgpu_set_gather(zk_mgpu* mg, int mode) {
    if (!mg || (mode != 0 && mode != 1)) return ZK_ERR_ARG;
    if (mode == 1 && mg->comms.empty()) {
        if (!mg->nccl.load()) { set_err(mg, "zk_mgpu_set_gather", "libnccl.so.2 not loadable"); return ZK_ERR_CUDA; }
        mg->comms.assign(mg->g, nullptr);
        ncclResult_t r = mg->nccl.CommInitAll(mg->comms.data(), mg->g, mg->devices.data());
        if (r != 0) {
            set_err(mg, "ncclCommInitAll", mg->nccl.GetErrorString ? mg->nccl.GetErrorString(r) : "failed");
            mg->comms.clear();
            return ZK_ERR_CUDA;
        }
        mg->gather_all.assign(mg->g, nullptr);
        for (int d = 0; d < mg->g; d++) {
            cudaSetDevice(mg->devices[d]);
            if (cudaMalloc((void**)&mg->gather_all[d], (size_t)mg->g * 128) != cudaSuccess) { cudaGetLastError(); return ZK_ERR_NOMEM; }
        }
    }
    mg->gather_mode = mode;
    return ZK_OK;
}

// Partials -> device 0 -> one 32-byte encoding.  Called after every device has QUEUED its MSM.
static int gather_and_encode(zk_mgpu* mg, uint8_t out32[32]) {
    const int g = mg->g;
    const void* src0 = nullptr;
    if (mg->gather_mode == 1) {
        ncclResult_t r = mg->nccl.GroupStart();
        for (int d = 0; d < g && r == 0; d++)
            r = mg->nccl.AllGather(zk_internal_partial_ptr(mg->ctx[d]), mg->gather_all[d], 128, kNcclUint8, mg->comms[d], mg->ctx[d]->stream);
        ncclResult_t r2 = mg->nccl.GroupEnd();
        if (r != 0 || r2 != 0) { set_err(mg, "ncclAllGather", mg->nccl.GetErrorString ? mg->nccl.GetErrorString(r ? r : r2) : "failed"); return ZK_ERR_CUDA; }
        src0 = mg->gather_all[0];               // ordered on device 0's stream by NCCL itself
    } else {
        for (int d = 0; d < g; d++) {
            zk_ctx* c = mg->ctx[d];
            CK(c, cudaSetDevice(mg->devices[d]));
            if (d == 0) CK(c, cudaMemcpyAsync(mg->gather0, zk_internal_partial_ptr(c), 128, cudaMemcpyDeviceToDevice, c->stream));
            else CK(c, cudaMemcpyPeerAsync(mg->gather0 + (size_t)d * 128, mg->devices[0], zk_internal_partial_ptr(c), mg->devices[d], 128, c->stream));
            CK(c, cudaEventRecord(mg->ev_part[d], c->stream));
        }
        zk_ctx* c0 = mg->ctx[0];
        CK(c0, cudaSetDevice(mg->devices[0]));
        for (int d = 1; d < g; d++) CK(c0, cudaStreamWaitEvent(c0->stream, mg->ev_part[d], 0));
        src0 = mg->gather0;
    }
    return zk_ext_sum_compress_dev(mg->ctx[0], src0, (size_t)g, out32);
}

// wait for every device; the lowest rejected GLOBAL index wins
static int finish_all(zk_mgpu* mg, const std::vector<size_t>& global_base, size_t* bad_index) {
    int rc = ZK_OK; size_t best = (size_t)-1;
    for (int d = 0; d < mg->g; d++) {
        size_t b = (size_t)-1;
        int x = zk_internal_finish_partial(mg->ctx[d], &b);
        if (x == ZK_ERR_INVALID_POINT) { if (global_base[d] + b < best) best = global_base[d] + b; if (rc == ZK_OK) rc = x; }
        else if (x != ZK_OK) rc = x;
    }
    if (bad_index && best != (size_t)-1) *bad_index = best;
    return rc;
}

extern "C" int zk_mgpu_msm_vartime(zk_mgpu* mg, const uint8_t* scalars32_host, const uint8_t* points32_host, size_t n, uint8_t out32[32]) {
    if (!mg || !out32 || (n && (!scalars32_host || !points32_host))) return ZK_ERR_ARG;
    mg->err[0] = 0;
    std::vector<size_t> base(mg->g, 0);
    int rc = run_all(mg, [&](int r) {
        size_t lo, hi; shard_range(n, r, mg->g, &lo, &hi);
        base[r] = lo;
        return zk_internal_enqueue_partial(mg->ctx[r], nullptr, 0, nullptr, 0, 0, scalars32_host + lo * 32, points32_host + lo * 32, hi - lo);
    });
    int rc2 = rc == ZK_OK ? gather_and_encode(mg, out32) : rc;
    int rc3 = finish_all(mg, base, nullptr);
    if (rc2 != ZK_OK) return rc2;
    if (rc3 != ZK_OK) { memset(out32, 0, 32); return rc3; }
    return ZK_OK;
}

// ---- sharded point cache ----
extern "C" int zk_mgpu_table_create(zk_mgpu* mg, size_t capacity, zk_mgpu_table** out) {
    if (!mg || !out) return ZK_ERR_ARG;
    *out = nullptr;
    zk_mgpu_table* t = new (std::nothrow) zk_mgpu_table();
    if (!t) return ZK_ERR_NOMEM;
    t->mg = mg;
    for (int r = 0; r < mg->g; r++) {
        zk_table* s = nullptr;
        int rc = zk_table_create(mg->ctx[r], (capacity + mg->g - 1) / mg->g, &s);
        if (rc != ZK_OK) { for (zk_table* x : t->shard) zk_table_destroy(x); delete t; return rc; }
        t->shard.push_back(s);
    }
    *out = t;
    return ZK_OK;
}
extern "C" void zk_mgpu_table_destroy(zk_mgpu_table* t) {
    if (!t) return;
    for (zk_table* s : t->shard) zk_table_destroy(s);
    delete t;
}
extern "C" size_t zk_mgpu_table_len(const zk_mgpu_table* t) { return t ? t->len : 0; }
extern "C" void zk_mgpu_table_clear(zk_mgpu_table* t) {
    if (!t) return;
    for (zk_table* s : t->shard) zk_table_clear(s);
    t->segs.clear(); t->len = 0;
}

// item_bytes = 32 (compressed encodings) or 64 (uniform strings for hash-to-group)
static int mgpu_append(zk_mgpu_table* t, const uint8_t* host, size_t n, size_t item_bytes, size_t* bad_index) {
    if (!t || (!host && n)) return ZK_ERR_ARG;
    zk_mgpu* mg = t->mg;
    mg->err[0] = 0;
    if (n == 0) return ZK_OK;
    std::vector<size_t> bad(mg->g, (size_t)-1), lo_of(mg->g, 0);
    int rc = run_all(mg, [&](int r) {
        size_t lo, hi; shard_range(n, r, mg->g, &lo, &hi);
        lo_of[r] = lo;
        if (item_bytes == 32) return zk_table_append_compressed(mg->ctx[r], t->shard[r], host + lo * 32, hi - lo, &bad[r]);
        return zk_table_append_uniform(mg->ctx[r], t->shard[r], host + lo * 64, hi - lo);
    });
    if (rc != ZK_OK) {
        // all-or-nothing: roll the shards that did append back to their previous length
        size_t best = (size_t)-1;
        for (int r = 0; r < mg->g; r++) {
            size_t lo, hi; shard_range(n, r, mg->g, &lo, &hi);
            size_t want = 0;
            for (const auto& s : t->segs) { size_t a, b; shard_range(s.n, r, mg->g, &a, &b); want += b - a; }
            if (zk_table_len(t->shard[r]) > want) t->shard[r]->len = want;
            if (bad[r] != (size_t)-1 && lo_of[r] + bad[r] < best) best = lo_of[r] + bad[r];
        }
        if (bad_index && best != (size_t)-1) *bad_index = best;
        return rc;
    }
    t->segs.push_back({t->len, n});
    t->len += n;
    return ZK_OK;
}
extern "C" int zk_mgpu_table_append_compressed(zk_mgpu_table* t, const uint8_t* points32_host, size_t n, size_t* bad_index) {
    return mgpu_append(t, points32_host, n, 32, bad_index);
}
extern "C" int zk_mgpu_table_append_uniform(zk_mgpu_table* t, const uint8_t* bytes64_host, size_t n) {
    return mgpu_append(t, bytes64_host, n, 64, nullptr);
}

// static part: table[offset .. offset+n_static) with scalars_static; dynamic part: n_dyn compressed encodings with their
// scalars.  Device r takes its own rows of the slice plus the index range shard_range(n_dyn, r, g) of the dynamic terms.
extern "C" int zk_mgpu_msm_vartime_mixed(zk_mgpu* mg, const uint8_t* scalars_static32_host, const zk_mgpu_table* t, size_t offset,
                                         size_t n_static, const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host,
                                         size_t n_dyn, uint8_t out32[32]) {
    if (!mg || !out32) return ZK_ERR_ARG;
    if (n_static && (!t || t->mg != mg || !scalars_static32_host || offset > t->len || n_static > t->len - offset)) return ZK_ERR_ARG;
    if (n_dyn && (!scalars_dyn32_host || !points_dyn32_host)) return ZK_ERR_ARG;
    mg->err[0] = 0;
    // device r: which of its local rows fall into [offset, offset+n_static), and which host scalar pieces feed them
    std::vector<std::vector<zk_host_piece>> pieces(mg->g);
    std::vector<size_t> row0(mg->g, 0), rows(mg->g, 0), dyn_lo(mg->g, 0), dyn_hi(mg->g, 0);
    for (int r = 0; r < mg->g; r++) {
        size_t local = 0; bool first = true;
        if (n_static) for (const auto& s : t->segs) {
            size_t a, b; shard_range(s.n, r, mg->g, &a, &b);
            size_t glo = s.base + a, ghi = s.base + b;                 // global indices of this device's rows of append s
            size_t lo = glo > offset ? glo : offset, hi = ghi < offset + n_static ? ghi : offset + n_static;
            if (lo < hi) {
                if (first) { row0[r] = local + (lo - glo); first = false; }
                pieces[r].push_back({scalars_static32_host + (lo - offset) * 32, (hi - lo) * 32});
                rows[r] += hi - lo;
            }
            local += b - a;
        }
        shard_range(n_dyn, r, mg->g, &dyn_lo[r], &dyn_hi[r]);
    }
    int rc = run_all(mg, [&](int r) {
        const size_t nd = dyn_hi[r] - dyn_lo[r];
        return zk_internal_enqueue_partial(mg->ctx[r], pieces[r].data(), (int)pieces[r].size(), rows[r] ? t->shard[r] : nullptr, row0[r], rows[r],
                                           nd ? scalars_dyn32_host + dyn_lo[r] * 32 : nullptr, nd ? points_dyn32_host + dyn_lo[r] * 32 : nullptr, nd);
    });
    int rc2 = rc == ZK_OK ? gather_and_encode(mg, out32) : rc;
    int rc3 = finish_all(mg, dyn_lo, nullptr);
    if (rc2 != ZK_OK) return rc2;
    if (rc3 != ZK_OK) { memset(out32, 0, 32); return rc3; }
    return ZK_OK;
}

extern "C" int zk_mgpu_msm_vartime_table(zk_mgpu* mg, const uint8_t* scalars32_host, const zk_mgpu_table* t, size_t offset, size_t n,
                                         uint8_t out32[32]) {
    if (!t) return ZK_ERR_ARG;
    return zk_mgpu_msm_vartime_mixed(mg, scalars32_host, t, offset, n, nullptr, nullptr, 0, out32);
}
