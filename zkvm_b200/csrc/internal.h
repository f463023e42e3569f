// Host-side internals shared by msm.cu (single-device pipeline + C ABI) and mgpu.cu (single-process multi-GPU layer).
// Nothing here crosses the ABI: include/zkmsm.h only exposes the opaque handles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/zkmsm.h"

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
};

// Pinned staging ring for uploads from PAGEABLE host memory (a Rust Vec<u8>, a numpy array): the calling thread
// memcpy()s chunk i+1 into a pinned slot while the DMA engine moves chunk i and the decoder works on chunk i-1.
// cudaMemcpyAsync from pageable memory would instead block the caller for the whole transfer and serialise with it.
constexpr int ZK_STAGE_SLOTS = 4;
constexpr size_t ZK_STAGE_SLOT_BYTES = (size_t)4 << 20;
constexpr size_t ZK_STAGE_MIN_BYTES = (size_t)256 << 10;   // smaller pageable copies go straight to cudaMemcpyAsync

struct zk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    char err[256] = {0};
    int forced_window = 0;
    int profiling = 0;
    float phase_ms[4] = {0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // copy/decode side streams: compressed points are uploaded and decoded in chunks on these two while the
    // main stream uploads the scalars and sorts digits; the main stream joins them right before the accumulation
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
    cudaEvent_t ev_scan = nullptr;          // histogram + scan done: a decoder that scatters its own terms may start
    bool join_aux = false;
    // high-priority stream for the short, latency-bound phases (digit sort, reduction tree, window Horner, Encode): with
    // several contexts in flight their grids must not queue behind the other contexts' bulk kernels (accumulate, decode),
    // which stay on the low-priority main / side streams.  The phases hop between the two through these events; the main
    // stream always waits for the hop back, so callers see one stream (msm.cu: ZK_TAIL_HP).
    cudaStream_t tail = nullptr;
    cudaEvent_t ev_acc = nullptr, ev_tail = nullptr, ev_pre = nullptr, ev_sort = nullptr;
    uint64_t launches = 0;
    int sm_count = 0;
    // workspace
    DevBuf scalars, comp, dyn_table, counts, cursor, offsets, tiles, entries, partials, task_off, tasks, plan, tree_a, tree_w,
        out_ext, out32, bad, seg, batch_ext, batch_out, inv_scratch;
    uint8_t* h_out = nullptr;               // pinned 64 B: [0,32) encoding, [32,40) bad index
    // staging ring (allocated on first use)
    uint8_t* stage[ZK_STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[ZK_STAGE_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_busy[ZK_STAGE_SLOTS] = {false, false, false, false};
    unsigned stage_next = 0;
    int staging_mode = 0;                   // 0 = auto (ring for pageable sources), 1 = never, 2 = always
    uint64_t staged_bytes = 0;              // bytes that went through the ring (diagnostics)
    int wait_mode = 0;                      // 0 = spin in cudaStreamSynchronize, 1 = park on a blocking-sync event
    cudaEvent_t ev_done = nullptr;          // blocking-sync event for wait_main()
};

struct zk_table {
    int device = 0;
    uint4* d = nullptr;
    size_t len = 0, cap = 0;
    // optional window expansion (zk_table_precompute): pre[(w*pre_len + i)] = 2^(off_w) * point i, w < pre_W
    uint4* pre = nullptr;
    size_t pre_len = 0;
    int pre_c = 0, pre_W = 0;
};

#define CK(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return e_ == cudaErrorMemoryAllocation ? ZK_ERR_NOMEM : ZK_ERR_CUDA;                    \
        }                                                                                          \
    } while (0)
#define TRY(x) do { int rc_ = (x); if (rc_ != ZK_OK) return rc_; } while (0)

// ---- internal entry points (msm.cu) used by the multi-GPU layer ----
// One piece of a scalar upload: `bytes` bytes at `host` go to consecutive positions of the ctx's scalar buffer.
struct zk_host_piece { const uint8_t* host; size_t bytes; };

// Queue (asynchronously, on the ctx's streams) one MSM whose result stays on the device as a 128-byte extended point
// at zk_internal_partial_ptr(ctx).  Scalars come from `npieces` host pieces (concatenated); points are
// table[offset .. offset+n_static) followed by n_dyn compressed encodings at points_dyn32_host.  The caller must
// zk_internal_finish_partial() before reading anything back.
int zk_internal_enqueue_partial(zk_ctx* ctx, const zk_host_piece* static_pieces, int npieces, const zk_table* t, size_t offset,
                                size_t n_static, const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host, size_t n_dyn);
void* zk_internal_partial_ptr(zk_ctx* ctx);
// Wait for the ctx's stream; *bad_index = lowest rejected dynamic encoding (SIZE_MAX when none); returns
// ZK_ERR_INVALID_POINT in that case.
int zk_internal_finish_partial(zk_ctx* ctx, size_t* bad_index);
