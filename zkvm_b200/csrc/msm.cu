// Ristretto255 vartime MSM for sm_100a: kernels, the per-context pipeline and the C ABI (include/zkmsm.h).
//
// Pipeline (one MSM = one pass; everything stays in HBM between kernels):
//   decompress      32 B encodings -> 96 B affine-Niels table entries   (IMAD-bound, ~265 field ops/point)
//   digits+hist     scalars mod l -> signed radix-2^c digits -> per-(window,bucket) counts (L2 atomics)
//   scan            exclusive prefix sum over the W*2^(c-1) counts
//   scatter         point index (+ sign bit) written at its bucket's cursor  = counting sort by bucket
//   bucket accum    one thread per bucket walks its sorted index run, 7M mixed adds against the Niels table
//   tree reduce     radix-8 tree per window: (A, Wt) = (plain sum, position-weighted sum) of bucket ranges
//   window combine  Horner over windows + RFC 9496 Encode, one launch (k_combine_encode)
// Host side (second half of this file): contexts and workspaces, host->device staging, the pipeline driver and the
// single-device C ABI; the multi-GPU entry points live in mgpu.cu.
//
// Spec: RFC 9496 for every byte that crosses the ABI; the bucket method itself is the textbook
// Pippenger/Bernstein algorithm (the result is algorithm-independent: the encoding of a group
// element is canonical).  No reference source is mounted.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>

#include "../../include/zkmsm.h"
#include "ge25519.cuh"
#include "recode.cuh"

using namespace zk;

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int REDUCE_RADIX_LOG2 = 3;               // tree fan-in 8
constexpr int REDUCE_RADIX = 1 << REDUCE_RADIX_LOG2;
#ifndef ZK_TREE_WARP_MAX
#define ZK_TREE_WARP_MAX 2048
#endif
constexpr size_t TREE_WARP_MAX_NODES = ZK_TREE_WARP_MAX;   // upper tree levels with at most this many nodes use a warp per node
constexpr uint32_t TASK_LEN = 64;                   // longest run of entries one accumulation thread walks
// CTA sizes of the sort / scan / tree kernels (tunables; co-scheduling them beside another context's accumulation was
// measured and does not happen: the hardware starts CTAs of a second grid only once the first has none left to dispatch).
#ifndef ZK_SORT_BLOCK
#define ZK_SORT_BLOCK 256
#endif
#ifndef ZK_TREE_BLOCK
#define ZK_TREE_BLOCK 128
#endif
#ifndef ZK_SCAN_BLOCK
#define ZK_SCAN_BLOCK 256
#endif
#ifndef ZK_SCATTER_BATCH
#define ZK_SCATTER_BATCH 8                          // returning atomics a scatter thread keeps in flight
#endif
// Stream priorities inside one context (several contexts are in flight per device):
//   0 = everything on the context's one stream;
//   1 = the phases after the accumulation (tree, Horner, Encode) on a high-priority stream;
//   2 = the digit sort too: only the bulk kernels (accumulate, decode) stay at low priority, so the short, latency-bound
//       phases of one MSM are never queued behind the multi-wave grids of the others.
#ifndef ZK_TAIL_HP
#define ZK_TAIL_HP 2
#endif
constexpr int SCAN_BLOCK = ZK_SCAN_BLOCK;           // threads per scan CTA, 4 buckets each
constexpr int SCAN_TILE = SCAN_BLOCK * 4;

static inline int windows_for_width(int c) { return (254 + c - 1) / c; }           // fewest windows of width <= c
static inline int max_width(int W) { return (254 + W - 1) / W; }

__device__ __forceinline__ void ld_fe(fe& r, const uint4* p) {
    uint4 a = __ldg(p), b = __ldg(p + 1);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
}
__device__ __forceinline__ void ld_fe_plain(fe& r, const uint4* p) {   // data written earlier in this stream by another kernel
    uint4 a = p[0], b = p[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
}
__device__ __forceinline__ void st_fe(uint4* p, const fe& r) {
    p[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    p[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void ld_niels(ge_niels& q, const uint4* base, size_t idx) {
    const uint4* p = base + idx * 6;
    ld_fe(q.yp, p); ld_fe(q.ym, p + 2); ld_fe(q.t2d, p + 4);
}
__device__ __forceinline__ void st_niels(uint4* base, size_t idx, const ge_niels& q) {
    uint4* p = base + idx * 6;
    st_fe(p, q.yp); st_fe(p + 2, q.ym); st_fe(p + 4, q.t2d);
}
__device__ __forceinline__ void ld_ext(ge_ext& r, const uint4* base, size_t idx) {
    const uint4* p = base + idx * 8;
    ld_fe_plain(r.X, p); ld_fe_plain(r.Y, p + 2); ld_fe_plain(r.Z, p + 4); ld_fe_plain(r.T, p + 6);
}
__device__ __forceinline__ void st_ext(uint4* base, size_t idx, const ge_ext& r) {
    uint4* p = base + idx * 8;
    st_fe(p, r.X); st_fe(p + 2, r.Y); st_fe(p + 4, r.Z); st_fe(p + 6, r.T);
}

// Batch mode: `seg` (M+1 ascending offsets, seg[0] = 0, seg[M] = n) splits the n terms into M independent MSMs.
// MSM m owns windows [m*W, (m+1)*W) of the bucket space; everything downstream only sees "M*W windows".
__device__ __forceinline__ uint32_t segment_of(const uint32_t* __restrict__ seg, uint32_t M, uint32_t i) {
    uint32_t lo = 0, hi = M;                 // invariant: seg[lo] <= i < seg[hi]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (seg[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

// ---- point-format kernels --------------------------------------------------------------------

// RFC 9496 4.3.1 over a batch; writes affine-Niels entries.  *bad = lowest rejected index.
#ifndef ZK_DEC_PAIR
#define ZK_DEC_PAIR 0   // measured: 2.16 ms vs 1.98 ms at 2^20 (205 registers halve the occupancy); kept for the record
#endif
// One thread: encoding i of `in` -> affine-Niels row i of `table` (identity + reject flags when it does not decode).
__device__ __forceinline__ void decompress_one(const uint4* __restrict__ in, size_t i, uint4* __restrict__ table,
                                               unsigned long long* __restrict__ bad, unsigned long long index_base,
                                               const uint32_t* __restrict__ seg, uint32_t M, uint32_t* __restrict__ bad_msm) {
    uint4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    fe x, y, t;
    bool ok = ristretto_decode(x, y, t, w);
    ge_niels q;
    if (ok) ge_to_niels_affine(q, x, y, t);
    else {
        ge_niels_identity(q);
        atomicMin(bad, index_base + (unsigned long long)i);
        if (bad_msm) bad_msm[segment_of(seg, M, (uint32_t)(index_base + i))] = 1u;    // batch mode: only that MSM is void
    }
    st_niels(table, i, q);
}
// Each thread decodes points 2t and 2t+1 with one interleaved exponentiation chain (ZK_DEC_PAIR), see fe2.
__global__ void __launch_bounds__(128) k_decompress(const uint4* __restrict__ in, size_t n, uint4* __restrict__ table,
                                                    unsigned long long* __restrict__ bad, unsigned long long index_base,
                                                    const uint32_t* __restrict__ seg = nullptr, uint32_t M = 0,
                                                    uint32_t* __restrict__ bad_msm = nullptr) {
#if ZK_DEC_PAIR
    size_t i0 = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i0 >= n) return;
    size_t i1 = i0 + 1 < n ? i0 + 1 : i0;            // odd tail: decode the last point twice, store it once
    uint4 a0 = __ldg(in + 2 * i0), b0 = __ldg(in + 2 * i0 + 1), a1 = __ldg(in + 2 * i1), b1 = __ldg(in + 2 * i1 + 1);
    uint32_t w0[8] = {a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w};
    uint32_t w1[8] = {a1.x, a1.y, a1.z, a1.w, b1.x, b1.y, b1.z, b1.w};
    fe x[2], y[2], t[2]; bool ok[2];
    ristretto_decode_x2(x[0], y[0], t[0], ok[0], w0, x[1], y[1], t[1], ok[1], w1);
#pragma unroll
    for (int k = 0; k < 2; k++) {
        size_t i = k ? i1 : i0;
        if (k && i1 == i0) break;
        ge_niels q;
        if (ok[k]) ge_to_niels_affine(q, x[k], y[k], t[k]);
        else {
            ge_niels_identity(q);
            atomicMin(bad, index_base + (unsigned long long)i);
            if (bad_msm) bad_msm[segment_of(seg, M, (uint32_t)(index_base + i))] = 1u;    // batch mode: only that MSM is void
        }
        st_niels(table, i, q);
    }
#else
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    decompress_one(in, i, table, bad, index_base, seg, M, bad_msm);
#endif
}

// RFC 9496 4.3.4 over a batch of 64-byte strings -> extended points (any Z); k_ext_to_niels normalises them with a
// batched inversion afterwards (one exponentiation per INV_BATCH points instead of one per point).
__global__ void __launch_bounds__(128) k_map_uniform(const uint4* __restrict__ in, size_t n, uint4* __restrict__ ext_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 a = __ldg(in + 4 * i + k);
        w[4 * k] = a.x; w[4 * k + 1] = a.y; w[4 * k + 2] = a.z; w[4 * k + 3] = a.w;
    }
    ge_ext p;
    ristretto_from_uniform(p, w);
    st_ext(ext_out, i, p);
}

// Extended points (X, Y, Z, T as 4 x 32-byte little-endian field elements; any representative, any Z != 0) ->
// affine-Niels entries, fully validated.  This is how a caller that already holds decompressed points (dalek
// `RistrettoPoint` = four field elements) hands them over without compressing them on the CPU.  Rejected (index
// reported): a non-canonical field encoding, Z = 0, a point off the curve -x^2 + y^2 = 1 + d x^2 y^2, T*Z != X*Y,
// or a point outside the even subgroup 2E that ristretto255 representatives live in (RFC 9496 section 3: the group is
// 2E / E[4]).  Membership test: P in 2E  <=>  Z^2 - Y^2 is a square (checked exhaustively against the big-integer
// oracle in tests/test_oracle.py); the same inverse square root also yields 1/Z, so validation costs no extra
// exponentiation over the normalisation.
__global__ void __launch_bounds__(128) k_from_extended(const uint4* __restrict__ in, size_t n, uint4* __restrict__ table,
                                                       unsigned long long* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe c[4]; bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint4 a = __ldg(in + 8 * i + 2 * k), b = __ldg(in + 8 * i + 2 * k + 1);
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        ok &= fe_from_words(c[k], w);
    }
    fe zz, yy, u, wv, r, zi, x, y, t, xx, lhs, rhs, tmp, one = fe_one(), dd = fe_d();
    fe_sqr(zz, c[2]); fe_sqr(yy, c[1]);
    fe_sub(u, zz, yy);                                              // Z^2 - Y^2
    const bool z_zero = fe_is_zero(c[2]);
    const bool u_zero = fe_is_zero(u);
    fe_mul(wv, u, zz);
    bool sq = fe_invsqrt(r, wv);                                    // r^2 = 1 / (u Z^2) when that is a square
    fe_sqr(zi, r); fe_mul(zi, zi, u); fe_mul(zi, zi, c[2]);         // Z * r^2 * u = 1/Z
    fe_mul(x, c[0], zi); fe_mul(y, c[1], zi);
    // y = +-1 (u == 0, so x must be 0): the square-root route degenerates; y = Y/Z is +1 or -1 by comparison
    fe my; fe_neg(my, one);
    fe_select(tmp, my, one, fe_eq(c[1], c[2]));
    fe_select(y, y, tmp, u_zero);
    fe zero = fe_zero();
    fe_select(x, x, zero, u_zero);
    ok &= !z_zero & (sq | u_zero);
    ok &= !u_zero | fe_is_zero(c[0]);
    fe_mul(t, x, y);
    fe_mul(tmp, c[3], zi); fe_select(tmp, tmp, zero, u_zero);
    ok &= fe_eq(tmp, t);                                             // T/Z == (X/Z)(Y/Z)
    ok &= !u_zero | fe_is_zero(c[3]);
    fe_sqr(xx, x); fe_sqr(yy, y);
    fe_sub(lhs, yy, xx);
    fe_mul(rhs, xx, yy); fe_mul(rhs, rhs, dd); fe_add(rhs, rhs, one);
    ok &= fe_eq(lhs, rhs);
    ge_niels q;
    if (ok) ge_to_niels_affine(q, x, y, t); else { ge_niels_identity(q); atomicMin(bad, (unsigned long long)i); }
    st_niels(table, i, q);
}

__device__ __forceinline__ void niels_to_ext(ge_ext& p, const ge_niels& q) {
    // (2x : 2y : 2 : 2xy);  T = X*Y/Z = X*Y*2^-1
    fe inv2 = {{0xfffffff7u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu}};
    fe_sub(p.X, q.yp, q.ym); fe_add(p.Y, q.yp, q.ym);
    p.Z = fe_zero(); p.Z.v[0] = 2;
    fe_mul(p.T, p.X, p.Y); fe_mul(p.T, p.T, inv2);
}

// Window expansion of a static table: row w*len + i of the expanded table is the affine-Niels form of 2^(off_w) * P_i.
// Pass A walks the doubling chain of each point and parks the multiples in extended form; pass B normalises each to
// Z = 1 (one inversion per entry; this runs once per generator set).
__global__ void __launch_bounds__(128) k_precomp_double(const uint4* __restrict__ table, size_t len, int W, uint4* __restrict__ scratch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    ge_niels q; ld_niels(q, table, i);
    ge_ext r; niels_to_ext(r, q);
#pragma unroll 1
    for (int w = 1; w < W; w++) {
#pragma unroll 1
        for (int d = window_geom(W, w - 1).width; d > 0; d--) ge_dbl(r, r);
        st_ext(scratch, (size_t)(w - 1) * len + i, r);
    }
}
// Extended points (any Z != 0) -> affine-Niels entries with Montgomery's batched inversion: a thread owns INV_BATCH
// points (strided by the thread count, so every load/store is coalesced), multiplies their Z's together, inverts the
// product once (one ~265-operation exponentiation) and peels the individual inverses off with 3 multiplies per point.
// Also the whole of zk_table_append_extended_unchecked: caller-provided coordinates are taken modulo p (any 256-bit
// value is a loose representative) and only Z = 0 is rejected (lowest index -> *bad; that entry becomes the identity).
constexpr int INV_BATCH = 8;
__global__ void __launch_bounds__(128) k_ext_to_niels(const uint4* __restrict__ ext, size_t n, uint4* __restrict__ out,
                                                      unsigned long long* __restrict__ bad) {
    const size_t T = (n + INV_BATCH - 1) / INV_BATCH;               // threads that own at least one point
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    fe pre[INV_BATCH], z, acc = fe_one();
#pragma unroll
    for (int k = 0; k < INV_BATCH; k++) {
        const size_t i = (size_t)k * T + t;
        if (i < n) {
            ld_fe_plain(z, ext + i * 8 + 4);
            if (fe_is_zero(z)) { z = fe_one(); if (bad) atomicMin(bad, (unsigned long long)i); }
            fe_mul(acc, acc, z);
        }
        pre[k] = acc;
    }
    fe inv; fe_invert(inv, acc);
#pragma unroll
    for (int k = INV_BATCH - 1; k >= 0; k--) {
        const size_t i = (size_t)k * T + t;
        if (i >= n) continue;
        ge_ext p; ld_ext(p, ext, i);
        const bool zz = fe_is_zero(p.Z);
        if (zz) p.Z = fe_one();
        fe zi, x, y, tt;
        if (k > 0) fe_mul(zi, inv, pre[k - 1]); else zi = inv;
        fe_mul(inv, inv, p.Z);
        fe_mul(x, p.X, zi); fe_mul(y, p.Y, zi); fe_mul(tt, x, y);
        ge_niels q;
        if (zz) ge_niels_identity(q); else ge_to_niels_affine(q, x, y, tt);
        st_niels(out, i, q);
    }
}

// RFC 9496 4.3.2 over table entries.
__global__ void __launch_bounds__(128) k_compress_table(const uint4* __restrict__ table, size_t n, uint4* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge_niels q; ld_niels(q, table, i);
    ge_ext p; niels_to_ext(p, q);
    uint32_t o[8]; ristretto_encode(o, p);
    out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
    out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---- scalar recoding ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_digit_hist(const uint4* __restrict__ scalars, size_t n, int c, int W,
                                                    const uint32_t* __restrict__ seg, uint32_t M, uint32_t win_stride,
                                                    uint32_t* __restrict__ counts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = __ldg(scalars + 2 * i), b = __ldg(scalars + 2 * i + 1);
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    scalar_reduce(s);
    uint32_t carry = 0; int B = 1 << (c - 1);
    // precomputed-window tables (win_stride != 0): all windows of an MSM share ONE bucket set
    const int wpm = win_stride ? 1 : W;                  // bucket-space windows per MSM
    size_t wbase = seg ? (size_t)segment_of(seg, M, (uint32_t)i) * wpm : 0;
    for (int w = 0; w < W; w++) {
        int d = next_digit(s, window_geom(W, w), carry);
        if (d != 0) atomicAdd(&counts[(wbase + (win_stride ? 0 : w)) * B + (abs(d) - 1)], 1u);
    }
}

// The digits of one (reduced) scalar go to their buckets' entry runs: entry = point index | sign.  ZK_SCATTER_BATCH
// windows at a time: their returning atomics are issued back to back, then the stores.
// Precomputed mode (win_stride != 0): window w of point p is row w*win_stride + p of the expanded table (= 2^(c*w) * P)
// and all windows share the bucket set `wbase`.
__device__ __forceinline__ void scatter_digits(const uint32_t* s, uint32_t pidx, size_t wbase, int c, int W, uint32_t win_stride,
                                               uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
    uint32_t carry = 0; const int B = 1 << (c - 1);
    for (int w0 = 0; w0 < W; w0 += ZK_SCATTER_BATCH) {
        uint32_t pos[ZK_SCATTER_BATCH], val[ZK_SCATTER_BATCH];
#pragma unroll
        for (int k = 0; k < ZK_SCATTER_BATCH; k++) {
            int w = w0 + k;
            int d = w < W ? next_digit(s, window_geom(W, w), carry) : 0;
            val[k] = d != 0 ? ((pidx + (uint32_t)w * win_stride) | (d < 0 ? 0x80000000u : 0u)) : 0xffffffffu;
            pos[k] = d != 0 ? atomicAdd(&cursor[(wbase + (win_stride ? 0 : w)) * B + (abs(d) - 1)], 1u) : 0u;
        }
#pragma unroll
        for (int k = 0; k < ZK_SCATTER_BATCH; k++) if (val[k] != 0xffffffffu) entries[pos[k]] = val[k];
    }
}

__global__ void __launch_bounds__(256) k_digit_scatter(const uint4* __restrict__ scalars, size_t n, int c, int W,
                                                       const uint32_t* __restrict__ seg, uint32_t M, int shared_points,
                                                       uint32_t win_stride, uint32_t* __restrict__ cursor,
                                                       uint32_t* __restrict__ entries) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = __ldg(scalars + 2 * i), b = __ldg(scalars + 2 * i + 1);
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    scalar_reduce(s);
    uint32_t m = seg ? segment_of(seg, M, (uint32_t)i) : 0u;
    const int wpm = win_stride ? 1 : W;
    // the point of term i: the i-th point, or (every MSM of the batch runs over the SAME points) the (i - seg[m])-th
    uint32_t pidx = shared_points ? (uint32_t)i - seg[m] : (uint32_t)i;
    scatter_digits(s, pidx, (size_t)m * wpm, c, W, win_stride, cursor, entries);
}

// Decode + scatter in one kernel, for points that arrive compressed together with their scalars (zk_msm_vartime, the
// dynamic suffix of zk_msm_vartime_mixed): thread i decodes encoding i of the chunk and then scatters the digits of
// that term's scalar.  The decoder is bound by the integer-multiply pipe and issues no memory traffic; the scatter is
// bound by the load/store unit (16 returning atomics + 16 scattered stores per term): inside one kernel the second
// hides under the first, which separate kernels of one stream cannot do (and kernels of different streams do not
// co-reside).  `scalars` = the scalars of the chunk's terms, `pidx_base` = point index of its
// first term.  The histogram and the scan must be complete (cursor = bucket offsets).
#if !ZK_DEC_PAIR
__global__ void __launch_bounds__(128) k_decompress_scatter(const uint4* __restrict__ in, size_t n, uint4* __restrict__ table,
                                                            unsigned long long* __restrict__ bad, unsigned long long index_base,
                                                            const uint4* __restrict__ scalars, uint32_t pidx_base, int c, int W,
                                                            uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    decompress_one(in, i, table, bad, index_base, nullptr, 0, nullptr);
    uint4 a = __ldg(scalars + 2 * i), b = __ldg(scalars + 2 * i + 1);
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    scalar_reduce(s);
    scatter_digits(s, pidx_base + (uint32_t)i, 0, c, W, 0u, cursor, entries);
}
#endif

// ---- scan + task planning (3 phases, SCAN_TILE-bucket tiles) -----------------------------------------
// Besides the exclusive scan of the bucket counts (-> offsets into `entries`), the same three passes split
// every bucket into tasks of at most TASK_LEN entries and emit the task list SORTED BY LENGTH (longest
// first).  One thread later runs one task, so the 32 lanes of a warp walk runs of (almost) equal length and
// the grid's tail is made of the shortest tasks: the bucket-size distribution (Poisson per window, 8x heavier
// in the top window, arbitrary for adversarial scalars) no longer decides the kernel's efficiency.
//   scanned value (packed u64): lo = entries in the bucket, hi = tasks of the bucket
//   bins[len], len in 1..TASK_LEN: number of tasks of that length; cursor = exclusive scan, descending length

__device__ __forceinline__ unsigned long long warp_incl_scan64(unsigned long long v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}
// exclusive scan of one u64 per thread across a SCAN_BLOCK-thread block; *total = block sum
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long* total) {
    __shared__ unsigned long long wsum[SCAN_BLOCK / 32];
    unsigned long long inc = warp_incl_scan64(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    unsigned long long off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_BLOCK / 32; k++) { unsigned long long x = wsum[k]; if (k < wid) off += x; tot += x; }
    *total = tot;
    __syncthreads();
    return off + inc - v;
}
__device__ __forceinline__ unsigned long long pack_count(uint32_t cnt) {
    return (unsigned long long)cnt | ((unsigned long long)((cnt + TASK_LEN - 1) / TASK_LEN) << 32);
}

// plan[0] = total tasks, plan[1 + len] = bins (phase 1) / cursors (phase 2 on), len = 0..TASK_LEN
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_tile_sums(const uint32_t* __restrict__ counts, size_t nb,
                                                        unsigned long long* __restrict__ tile_sums, uint32_t* __restrict__ plan) {
    __shared__ uint32_t s_bins[TASK_LEN + 1];
    for (int i = threadIdx.x; i <= TASK_LEN; i += SCAN_BLOCK) s_bins[i] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < nb) {
            uint32_t cnt = counts[base + k];
            v += pack_count(cnt);
            uint32_t nfull = cnt / TASK_LEN, rem = cnt % TASK_LEN;
            if (nfull) atomicAdd(&s_bins[TASK_LEN], nfull);
            if (rem) atomicAdd(&s_bins[rem], 1u);
        }
    }
    unsigned long long tot; block_excl_scan(v, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
    for (int i = threadIdx.x; i <= TASK_LEN; i += SCAN_BLOCK) if (s_bins[i]) atomicAdd(&plan[1 + i], s_bins[i]);
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_tiles(unsigned long long* __restrict__ tile_sums, size_t ntiles, uint32_t* __restrict__ plan) {
    // single block
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (size_t base = 0; base < ntiles; base += SCAN_BLOCK) {
        size_t i = base + threadIdx.x;
        unsigned long long v = i < ntiles ? tile_sums[i] : 0ull, tot;
        unsigned long long ex = block_excl_scan(v, &tot);
        unsigned long long cb = carry_s;
        if (i < ntiles) tile_sums[i] = cb + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = cb + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int len = TASK_LEN; len >= 1; len--) { uint32_t c = plan[1 + len]; plan[1 + len] = run; run += c; }
        plan[0] = run;
    }
}
// writes offsets/cursor (entry positions), task_off (first partial slot of each bucket) and the sorted task list
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const uint32_t* __restrict__ counts, size_t nb,
                                                    const unsigned long long* __restrict__ tile_offs, uint32_t* __restrict__ offsets,
                                                    uint32_t* __restrict__ cursor, uint32_t* __restrict__ task_off,
                                                    uint32_t* __restrict__ plan, uint2* __restrict__ tasks) {
    __shared__ uint32_t s_bins[TASK_LEN + 1];
    __shared__ uint32_t s_base[TASK_LEN + 1];
    for (int i = threadIdx.x; i <= TASK_LEN; i += SCAN_BLOCK) s_bins[i] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    uint32_t cnt[4], r_full[4], r_rem[4];
    unsigned long long sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        cnt[k] = base + k < nb ? counts[base + k] : 0u;
        sum += pack_count(cnt[k]);
        uint32_t nfull = cnt[k] / TASK_LEN, rem = cnt[k] % TASK_LEN;
        r_full[k] = nfull ? atomicAdd(&s_bins[TASK_LEN], nfull) : 0u;
        r_rem[k] = rem ? atomicAdd(&s_bins[rem], 1u) : 0u;
    }
    unsigned long long tot;
    unsigned long long ex = block_excl_scan(sum, &tot) + tile_offs[blockIdx.x];   // has a __syncthreads
    for (int i = threadIdx.x; i <= TASK_LEN; i += SCAN_BLOCK) s_base[i] = s_bins[i] ? atomicAdd(&plan[1 + i], s_bins[i]) : 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < nb) {
            uint32_t b = (uint32_t)(base + k);
            offsets[b] = (uint32_t)ex; cursor[b] = (uint32_t)ex; task_off[b] = (uint32_t)(ex >> 32);
            uint32_t nfull = cnt[k] / TASK_LEN, rem = cnt[k] % TASK_LEN;
            uint32_t p = s_base[TASK_LEN] + r_full[k];
            for (uint32_t j = 0; j < nfull; j++) tasks[p + j] = make_uint2(b, j);
            if (rem) tasks[s_base[rem] + r_rem[k]] = make_uint2(b, nfull);
        }
        ex += pack_count(cnt[k]);
        if (base + k == nb - 1) { offsets[nb] = (uint32_t)ex; task_off[nb] = (uint32_t)(ex >> 32); }
    }
}

// ---- bucket accumulation ------------------------------------------------------------------------

// One thread per task = up to TASK_LEN consecutive entries of one bucket's sorted run.  An entry is a point index
// (bit 31 = subtract); index space [0, split) -> tab_a, the rest -> tab_b.  The partial sum goes to
// partials[task_off[bucket] + j]; the tree's leaf level adds a bucket's partials together.
#ifndef ZK_ACCUM_MINBLOCKS
#define ZK_ACCUM_MINBLOCKS 3
#endif
#ifndef ZK_ACCUM_L2PREFETCH
#define ZK_ACCUM_L2PREFETCH 1
#endif
__device__ __forceinline__ void ld_point(ge_niels& q, const uint4* __restrict__ tab_a, const uint4* __restrict__ tab_b, uint32_t split, uint32_t e) {
    uint32_t idx = e & 0x7fffffffu;
    if (idx < split) ld_niels(q, tab_a, idx); else ld_niels(q, tab_b, idx - split);
}
__global__ void __launch_bounds__(128, ZK_ACCUM_MINBLOCKS)
k_bucket_accum(const uint4* __restrict__ tab_a, const uint4* __restrict__ tab_b, uint32_t split,
               const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
               const uint32_t* __restrict__ task_off, const uint2* __restrict__ tasks,
               const uint32_t* __restrict__ plan, uint4* __restrict__ partials) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plan[0]) return;
    uint2 d = tasks[t];
    uint32_t lo = offsets[d.x] + d.y * TASK_LEN, end = offsets[d.x + 1];
    uint32_t hi = lo + TASK_LEN < end ? lo + TASK_LEN : end;
    ge_ext acc;
    {   // a task is never empty: start from its first point (1 multiply) instead of identity + point (7)
        uint32_t e = entries[lo];
        ge_niels q; ld_point(q, tab_a, tab_b, split, e);
        ge_from_niels(acc, q, (e >> 31) != 0);
    }
#if ZK_ACCUM_L2PREFETCH
    uint32_t e = lo + 1 < hi ? entries[lo + 1] : 0u;
#pragma unroll 1
    for (uint32_t k = lo + 1; k < hi; k++) {
        // next index one iteration ahead; its point is pulled towards the SM with prefetches (no registers held)
        uint32_t en = k + 1 < hi ? entries[k + 1] : e;
        {
            uint32_t idx = en & 0x7fffffffu;
            const uint4* pp = idx < split ? tab_a + (size_t)idx * 6 : tab_b + (size_t)(idx - split) * 6;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 4));
        }
        ge_niels q; ld_point(q, tab_a, tab_b, split, e);
        ge_madd(acc, acc, q, (e >> 31) != 0);
        e = en;
    }
#else
#pragma unroll 1
    for (uint32_t k = lo + 1; k < hi; k++) {
        uint32_t e = entries[k];
        ge_niels q; ld_point(q, tab_a, tab_b, split, e);
        ge_madd(acc, acc, q, (e >> 31) != 0);
    }
#endif
    // stored in cached form (Y-X, Y+X, 2Z, 2dT): the tree's leaf level adds every partial into a running sum, and an
    // addition whose second operand is cached costs two multiplication rounds instead of three
    {
        ge_ext cch; const fe d2 = fe_d2();
        fe_sub(cch.X, acc.Y, acc.X); fe_add(cch.Y, acc.Y, acc.X); fe_dbl(cch.Z, acc.Z); fe_mul(cch.T, acc.T, d2);
        st_ext(partials, task_off[d.x] + d.y, cch);
    }
}

// ---- radix-8 reduction tree ---------------------------------------------------------------------
// A node covering children i = 0..L-1 (each of width `wc` buckets) combines
//     A  = sum_i A_i                       (plain sum)
//     Wt = sum_i Wt_i + wc * sum_i i*A_i   (sum weighted by 1-based position inside the node)
// At the leaves A_i = Wt_i = bucket i (wc = 1), and bucket i is itself the sum of its tasks' partials
// (none for an empty bucket).  The root's Wt is the window sum  sum_b (b+1) * S_b.
// ---- 4-lane cooperative point arithmetic ("quads") ----------------------------------------------
// The reduction tree and the window Horner are chains of dependent point operations; register-hungry scalar
// code runs them at 2 warps per scheduler.  A quad = 4 adjacent lanes holding one extended point, lane q owning
// one coordinate (0:X 1:Y 2:Z 3:T).  The 4 independent field multiplies of each half of the HWCD formulas run in
// the 4 lanes at once (the layout dalek's AVX2 backend uses across vector lanes); operands move between lanes
// with warp shuffles restricted to the quad's own 4-lane mask, so quads of one warp may diverge from each other.
// Depth per addition: 3 multiplies instead of 9; per doubling: 1 squaring + 1 multiply instead of 4 + 4; ~70
// registers per thread.  Roles are chosen with selects, never with branches.
struct quad_ctx { int q, base; unsigned mask; };
__device__ __forceinline__ quad_ctx quad_self() {
    quad_ctx c; c.q = threadIdx.x & 3; c.base = (threadIdx.x & 31) & ~3; c.mask = 0xfu << c.base; return c;
}
__device__ __forceinline__ void fe_shfl(fe& r, const fe& a, int src, unsigned mask) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(mask, a.v[i], src);
}
__device__ __forceinline__ void fe_shfl_xor(fe& r, const fe& a, int m, unsigned mask) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(mask, a.v[i], m);
}
__device__ __forceinline__ void fe_shfl_down(fe& r, const fe& a, int d, unsigned mask) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(mask, a.v[i], d);
}
__device__ __forceinline__ void quad_identity(fe& r, int q) { r = fe_zero(); r.v[0] = (q == 1 || q == 2) ? 1u : 0u; }
// in: lane0 = E, lane1 = H, lane2 = F, lane3 = G.  out: (X3, Y3, Z3, T3) = (E*F, G*H, F*G, E*H) in lanes 0..3.
__device__ __forceinline__ void quad_finish(fe& r, const fe& val, const quad_ctx& c) {
    fe m1, m2, x;
    fe_shfl(m1, val, c.base + (c.q == 0 ? 2 : (c.q == 3 ? 0 : 3)), c.mask);
    fe_shfl(m2, val, c.base + 1, c.mask);
    fe_select(x, val, m2, c.q == 3);
    fe_mul(r, x, m1);
}
// The second operand of an addition in "cached" form: (Y2-X2 | Y2+X2 | 2 Z2 | 2d T2) in lanes 0..3.
__device__ __forceinline__ void quad_to_cached(fe& v, const fe& qq, const quad_ctx& c) {
    const int q = c.q;
    const bool l0 = q == 0, hi = q >= 2;
    fe oq, a, b, dif, sum;
    fe_shfl_xor(oq, qq, 1, c.mask);
    fe_select(a, qq, oq, l0); fe_select(b, oq, qq, l0);
    fe_sub(dif, a, b); fe_add(sum, a, b);
    fe_select(v, sum, dif, l0); fe_select(v, v, qq, hi);        // Y2-X2 | Y2+X2 | Z2 | T2
    fe k = fe_d2();
#pragma unroll
    for (int i = 0; i < 8; i++) k.v[i] = q == 3 ? k.v[i] : (i == 0 ? (q == 2 ? 2u : 1u) : 0u);
    fe_mul(v, v, k);                                            // ... | ... | 2 Z2 | 2d T2
}
// r = p + (the point whose cached form is v).  Two multiplication rounds.  Unified a = -1 addition, complete.
__device__ __forceinline__ void quad_add_cached(fe& r, const fe& p, const fe& v, const quad_ctx& c) {
    const int q = c.q;
    const bool l0 = q == 0, hi = q >= 2;
    fe op, a, b, dif, sum, u;
    fe_shfl_xor(op, p, 1, c.mask);
    fe_select(a, p, op, l0); fe_select(b, op, p, l0);          // lanes 0/1: a = Y1, b = X1
    fe_sub(dif, a, b); fe_add(sum, a, b);
    fe_select(u, sum, dif, l0); fe_select(u, u, p, hi);         // Y1-X1 | Y1+X1 | Z1 | T1
    fe r1, o;
    fe_mul(r1, u, v);                                           // A | B | D | C
    fe_shfl_xor(o, r1, 1, c.mask);
    fe_add(sum, r1, o);
    fe_select(a, r1, o, l0); fe_select(b, o, r1, l0);
    fe_sub(dif, a, b);                                          // lane0: B-A = E, lane2: D-C = F
    fe val; fe_select(val, dif, sum, (q & 1) != 0);             // lane1: B+A = H, lane3: D+C = G
    quad_finish(r, val, c);
}
// r = p + qq (both in quad layout): three multiplication rounds.
__device__ __forceinline__ void quad_add(fe& r, const fe& p, const fe& qq, const quad_ctx& c) {
    fe v; quad_to_cached(v, qq, c);
    quad_add_cached(r, p, v, c);
}
// r = 2p, for chains of doublings (window Horner, the tree's weight shifts).  Compared with the straightforward lane
// assignment (every lane squares its own coordinate, lane 3 squares X+Y, t/A/B broadcast, two fetches before the final
// product: 56 shuffles) this takes 32 shuffles and one exchange round fewer:
// The lane that squares (X+Y) is the one that then owns E, so it needs no broadcast of t; every lane owns one factor of
// its final product and fetches the other with ONE shuffle (E <- F, F <- G, G <- H, H <- E).  The price is that the
// products land permuted: X and Z stay in lanes 0 and 2, Y and T swap lanes 1 and 3 on every doubling.  SW = the layout
// on entry (false: X Y Z T, true: X T Z Y); the result has layout !SW.  T is not an input of a doubling.
// Single-warp latency (tools/ubench/quad_latency.cu): 1386 cycles against 1510 for the straightforward form; exchanging
// through shared memory instead of shuffles was measured too and is slower (1540).
template <bool SW>
__device__ __forceinline__ void quad_dbl_chain(fe& r, const fe& p, const quad_ctx& c) {
    const int q = c.q;
    const int lx = c.base, lz = c.base + 2, ly = c.base + (SW ? 3 : 1), lt = c.base + (SW ? 1 : 3);
    const bool isx = q == 0, isz = q == 2, isy = q == (SW ? 3 : 1), ist = q == (SW ? 1 : 3);
    fe got, s, opnd, sq, A, Bv, apb, bma, c2, L, R, val, oth, z = fe_zero();
    fe_shfl(got, p, isx ? ly : (ist ? lx : c.base + q), c.mask);     // X-holder gets Y, T-holder gets X
    fe_add(s, p, got);
    fe_select(opnd, p, s, isx); fe_select(opnd, opnd, got, ist);     // X+Y | Y | Z | X
    fe_sqr(sq, opnd);                                               // t | B | Z^2 | A
    fe_shfl(A, sq, lt, c.mask); fe_shfl(Bv, sq, ly, c.mask);
    fe_add(apb, A, Bv); fe_sub(bma, Bv, A); fe_add(c2, sq, sq);
    fe_select(L, bma, sq, isx); fe_select(L, L, z, isy);             // t | 0 | B-A | B-A
    fe_select(R, apb, c2, isz); fe_select(R, R, z, ist);             // A+B | A+B | 2Z^2 | 0
    fe_sub(val, L, R);                                              // E | H | F | G   (holders of X | Y | Z | T)
    fe_shfl(oth, val, isx ? lz : (isz ? lt : (ist ? ly : lx)), c.mask);   // E<-F, F<-G, G<-H, H<-E
    fe_mul(r, val, oth);                                            // X3 = EF | T3 = EH | Z3 = FG | Y3 = GH
}
// k doublings of the quad's point, standard layout in and out.
__device__ __forceinline__ void quad_dbl_n(fe& acc, int k, const quad_ctx& c) {
#pragma unroll 1
    for (; k >= 2; k -= 2) { quad_dbl_chain<false>(acc, acc, c); quad_dbl_chain<true>(acc, acc, c); }
    if (k == 1) {
        quad_dbl_chain<false>(acc, acc, c);
        fe sw; fe_shfl(sw, acc, c.base + (c.q == 1 ? 3 : (c.q == 3 ? 1 : c.q)), c.mask);      // Y and T back to lanes 1 and 3
        acc = sw;
    }
}

__device__ __forceinline__ void quad_neg(fe& r, const fe& p, int q) { fe_cneg(r, p, q == 0 || q == 3); }
__device__ __forceinline__ void quad_ld(fe& r, const uint4* base, size_t idx, int q) { ld_fe_plain(r, base + idx * 8 + q * 2); }
__device__ __forceinline__ void quad_st(uint4* base, size_t idx, int q, const fe& r) { st_fe(base + idx * 8 + q * 2, r); }

constexpr uint32_t HEAVY_PARTIALS = 24;   // a bucket with more partials than this is summed by the whole warp

// Leaf level: run += bucket `idx`, i.e. every partial sum the accumulation left for that bucket (none: nothing to add).
// The partials are stored in cached form (k_bucket_accum), so each costs a two-round addition straight into the running
// sum.  Called by all 32 lanes at the same loop trip (valid = false for a quad without a child).  A bucket that was
// split into many tasks -- adversarial scalars put up to n / TASK_LEN partials in ONE bucket -- is reduced by the 8
// quads of the warp together: strided partial sums, then a butterfly over quads.
__device__ __forceinline__ void quad_add_bucket(fe& run, const uint4* __restrict__ partials, const uint32_t* __restrict__ task_off,
                                                size_t idx, bool valid, const quad_ctx& c) {
    uint32_t p0 = 0, p1 = 0;
    if (valid) { p0 = task_off[idx]; p1 = task_off[idx + 1]; }
    const bool heavy = p1 - p0 > HEAVY_PARTIALS;
    fe tmp;
    if (!heavy) {
#pragma unroll 1
        for (uint32_t p = p0; p < p1; p++) { quad_ld(tmp, partials, p, c.q); quad_add_cached(run, run, tmp, c); }
    }
    unsigned hmask = __ballot_sync(0xffffffffu, heavy);          // warp-wide rendezvous
    const int lane = threadIdx.x & 31;
    quad_ctx full = c; full.mask = 0xffffffffu;
    while (hmask) {
        int src = __ffs(hmask) - 1; hmask &= ~(0xfu << (src & ~3));
        uint32_t q0 = __shfl_sync(0xffffffffu, p0, src), q1 = __shfl_sync(0xffffffffu, p1, src);
        fe part; quad_identity(part, c.q);
#pragma unroll 1
        for (uint32_t p = q0 + (lane >> 2); p < q1; p += 8) { quad_ld(tmp, partials, p, c.q); quad_add_cached(part, part, tmp, c); }
#pragma unroll 1
        for (int o = 16; o >= 4; o >>= 1) { fe_shfl_xor(tmp, part, o, 0xffffffffu); quad_add(part, part, tmp, full); }
        fe sum; quad_add(sum, run, part, c);                     // every quad computes, only the owner keeps
        if ((lane & ~3) == (src & ~3)) run = sum;
    }
}

// One quad per tree node.  A node covering children i = 0..7 (each of width `wc` buckets) combines
//     A  = sum_i A_i                       (plain sum)
//     Wt = sum_i Wt_i + wc * sum_i i*A_i   (sum weighted by 1-based position inside the node)
// At the leaves (task_off != nullptr) A_i = Wt_i = bucket i, itself the sum of its tasks' partials.  The root's Wt
// is the window sum  sum_b (b+1) * S_b.  Loop trip counts are warp-uniform; missing children are the identity.
#ifndef ZK_TREE_MINBLOCKS
#define ZK_TREE_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(128, ZK_TREE_MINBLOCKS) k_tree_level_quad(const uint4* __restrict__ a_in, const uint4* __restrict__ wt_in,
                                                         const uint32_t* __restrict__ task_off, size_t m_in,
                                                         size_t m_out, int windows, int log2_wc,
                                                         uint4* __restrict__ a_out, uint4* __restrict__ wt_out) {
    size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const quad_ctx c = quad_self();
    const int q = c.q;
    size_t t = gt >> 2;
    const bool active = t < m_out * (size_t)windows;
    size_t w = active ? t / m_out : 0, k = active ? t % m_out : 0;
    size_t first = k * REDUCE_RADIX;
    fe run, acc, wsum, tmp;
    quad_identity(run, q); quad_identity(acc, q); quad_identity(wsum, q);
    if (task_off != nullptr) {
        // leaf level: children are global bucket ids w*m_in + j, looked up through task_off
#pragma unroll 1
        for (int jj = REDUCE_RADIX - 1; jj >= 0; jj--) {
            size_t j = first + jj;
            bool valid = active && j < m_in;
            quad_add_bucket(run, a_in, task_off, w * m_in + j, valid, c);
            quad_add(acc, acc, run, c);
        }
    } else {
        const uint4* ain = a_in + w * m_in * 8;
        const uint4* win = wt_in + w * m_in * 8;
#pragma unroll 1
        for (int jj = REDUCE_RADIX - 1; jj >= 0; jj--) {
            size_t j = first + jj;
            bool valid = active && j < m_in;
            if (valid) quad_ld(tmp, ain, j, q); else quad_identity(tmp, q);
            quad_add(run, run, tmp, c);
            quad_add(acc, acc, run, c);
            if (valid) quad_ld(tmp, win, j, q); else quad_identity(tmp, q);
            quad_add(wsum, wsum, tmp, c);
        }
        quad_neg(tmp, run, q); quad_add(acc, acc, tmp, c);      // sum i*A_i
        quad_dbl_n(acc, log2_wc, c);
        quad_add(acc, acc, wsum, c);
    }
    if (active) { quad_st(a_out, t, q, run); quad_st(wt_out, t, q, acc); }
}

// Upper tree levels when there are few nodes: one WARP per node, quad j owning child j.  The 24 dependent additions
// of the serial recurrence become a 3-step suffix scan plus 3-step butterflies across the quads of the warp
// (8 + 3*level addition depths instead of 26 + 3*level).  Costs 8x the lanes, so it is only used when the level
// is latency-bound (see msm_pipeline).
__global__ void __launch_bounds__(128) k_tree_level_warp(const uint4* __restrict__ a_in, const uint4* __restrict__ wt_in, size_t m_in,
                                                         size_t m_out, int windows, int log2_wc,
                                                         uint4* __restrict__ a_out, uint4* __restrict__ wt_out) {
    const size_t node = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, quad = lane >> 2;
    quad_ctx c = quad_self(); c.mask = 0xffffffffu;        // every lane of the warp takes part in every step
    const int q = c.q;
    const bool active = node < m_out * (size_t)windows;    // warp-uniform
    size_t w = active ? node / m_out : 0, k = active ? node % m_out : 0;
    size_t j = k * REDUCE_RADIX + quad;
    const bool valid = active && j < m_in;
    fe S, R, Ws, tmp, tmp2, ident;
    quad_identity(ident, q);
    if (valid) { quad_ld(S, a_in + w * m_in * 8, j, q); quad_ld(Ws, wt_in + w * m_in * 8, j, q); }
    else { S = ident; Ws = ident; }
    // suffix scan over the 8 quads: S_j = sum_{i >= j} A_i
#pragma unroll 1
    for (int d = 1; d < 8; d <<= 1) {
        fe_shfl_down(tmp, S, 4 * d, 0xffffffffu);
        fe_select(tmp, ident, tmp, quad + d < 8);
        quad_add(S, S, tmp, c);
    }
    // butterflies: R = sum_j S_j = sum_i (i+1) A_i ; Ws = sum_i Wt_i
    R = S;
#pragma unroll 1
    for (int o = 4; o < 32; o <<= 1) {
        fe_shfl_xor(tmp, R, o, 0xffffffffu); fe_shfl_xor(tmp2, Ws, o, 0xffffffffu);
        quad_add(R, R, tmp, c); quad_add(Ws, Ws, tmp2, c);
    }
    fe run; fe_shfl(run, S, q, 0xffffffffu);               // S_0 = plain sum, from quad 0
    quad_neg(tmp, run, q); quad_add(R, R, tmp, c);         // sum_i i*A_i
    quad_dbl_n(R, log2_wc, c);
    quad_add(R, R, Ws, c);
    if (active && quad == 0) { quad_st(a_out, node, q, run); quad_st(wt_out, node, q, R); }
}

// Horner over the per-window sums: out[m] = sum_w 2^(off_w) * Wt[m*W + w].  One quad per MSM (253 dependent doublings).
// `geomW` = number of digit windows the offsets come from (windows == 1 with precomputed tables: nothing to double).
__global__ void __launch_bounds__(32) k_window_combine(const uint4* __restrict__ wt, int windows, int geomW, uint32_t nmsm,
                                                       uint4* __restrict__ out_ext) {
    const uint32_t m = blockIdx.x * 8 + (threadIdx.x >> 2);
    if (m >= nmsm) return;
    const quad_ctx c = quad_self();
    const uint4* base = wt + (size_t)m * windows * 8;
    fe acc, tmp;
    quad_ld(acc, base, windows - 1, c.q);
#pragma unroll 1
    for (int w = windows - 2; w >= 0; w--) {
        quad_dbl_n(acc, window_geom(geomW, w).width, c);
        quad_ld(tmp, base, w, c.q);
        quad_add(acc, acc, tmp, c);
    }
    quad_st(out_ext, m, c.q, acc);
}

// out32[m] = Encode(ext[m]), one thread per point
__global__ void __launch_bounds__(64) k_encode_batch(const uint4* __restrict__ ext, size_t m, uint4* __restrict__ out32) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    ge_ext p; ld_ext(p, ext, i);
    uint32_t o[8]; ristretto_encode(o, p);
    out32[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
    out32[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
}
__global__ void k_set_identity_batch(uint4* __restrict__ out_ext, size_t m) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    ge_ext p; ge_identity(p); st_ext(out_ext, i, p);
}

// out32 = Encode(sum of g extended points).  One thread; the straight-line field code goes through fe_ops_call so that
// it stays in the instruction cache (see fe25519.cuh).
__global__ void k_ext_sum_encode(const uint4* __restrict__ ext, size_t g, uint4* __restrict__ out32) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ge_ext acc, tmp;
    ld_ext(acc, ext, 0);
#pragma unroll 1
    for (size_t i = 1; i < g; i++) { ld_ext(tmp, ext, i); ge_add(acc, acc, tmp); }
    uint32_t o[8]; ristretto_encode_ops<fe_ops_call>(o, acc);
    out32[0] = make_uint4(o[0], o[1], o[2], o[3]);
    out32[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// out32 = Encode(sum_i Decode(in[i])), g <= 1024 encodings, ONE launch of one CTA: thread i decodes point i (and folds
// points i + 256, ... into it), a shared-memory tree adds the per-thread sums, thread 0 encodes.  *bad = lowest rejected
// index.  This is the combine step of a multi-process deployment (every process returns the encoding of its partial MSM).
__global__ void __launch_bounds__(256) k_sum_compressed(const uint4* __restrict__ in, uint32_t g, uint4* __restrict__ out32,
                                                        unsigned long long* __restrict__ bad) {
    __shared__ uint4 sh[256 * 8];
    const uint32_t t = threadIdx.x;
    ge_ext acc; ge_identity(acc);
    for (uint32_t i = t; i < g; i += 256) {
        uint4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
        uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        fe x, y, tt;
        if (ristretto_decode(x, y, tt, w)) {
            ge_niels q; ge_to_niels_affine(q, x, y, tt);
            ge_madd(acc, acc, q, false);
        } else atomicMin(bad, (unsigned long long)i);
    }
    st_ext(sh, t, acc);
    __syncthreads();
    for (uint32_t stride = 128; stride >= 1; stride >>= 1) {
        if (t < stride && t + stride < 256 && t + stride < g) {
            ge_ext p, q; ld_ext(p, sh, t); ld_ext(q, sh, t + stride);
            ge_add(p, p, q); st_ext(sh, t, p);
        }
        __syncthreads();
    }
    if (t == 0) {
        ge_ext p; ld_ext(p, sh, 0);
        uint32_t o[8]; ristretto_encode_ops<fe_ops_call>(o, p);
        out32[0] = make_uint4(o[0], o[1], o[2], o[3]);
        out32[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// Single-MSM tail in ONE launch: Horner over the window sums (one quad), then RFC 9496 Encode (one thread) -- the
// extended result is also left in out_ext for callers that want it.  Saves a kernel boundary on the serial tail.
__global__ void __launch_bounds__(32) k_combine_encode(const uint4* __restrict__ wt, int windows, int geomW,
                                                       uint4* __restrict__ out_ext, uint4* __restrict__ out32) {
    __shared__ uint4 sh[8];
    if (threadIdx.x < 4) {
        const quad_ctx c = quad_self();
        fe acc, tmp;
        quad_ld(acc, wt, windows - 1, c.q);
#pragma unroll 1
        for (int w = windows - 2; w >= 0; w--) {
            quad_dbl_n(acc, window_geom(geomW, w).width, c);
            quad_ld(tmp, wt, w, c.q);
            quad_add(acc, acc, tmp, c);
        }
        quad_st(sh, 0, c.q, acc);
        quad_st(out_ext, 0, c.q, acc);
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        ge_ext p; ld_ext(p, sh, 0);
        uint32_t o[8]; ristretto_encode_ops<fe_ops_call>(o, p);
        out32[0] = make_uint4(o[0], o[1], o[2], o[3]);
        out32[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ---- integer-pipe microbenchmarks ---------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bench_imad_wide(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = a + i;
    for (int it = 0; it < iters; it++) {
        // two independent 4-slot carry chains per statement, the shape fe_mul issues
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0; madc.hi.cc.u32 %1, %16, %17, %1; madc.lo.cc.u32 %2, %16, %17, %2; madc.hi.cc.u32 %3, %16, %17, %3;"
            "madc.lo.cc.u32 %4, %16, %17, %4; madc.hi.cc.u32 %5, %16, %17, %5; madc.lo.cc.u32 %6, %16, %17, %6; madc.hi.u32 %7, %16, %17, %7;"
            "mad.lo.cc.u32 %8, %17, %16, %8; madc.hi.cc.u32 %9, %17, %16, %9; madc.lo.cc.u32 %10, %17, %16, %10; madc.hi.cc.u32 %11, %17, %16, %11;"
            "madc.lo.cc.u32 %12, %17, %16, %12; madc.hi.cc.u32 %13, %17, %16, %13; madc.lo.cc.u32 %14, %17, %16, %14; madc.hi.u32 %15, %17, %16, %15;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
              "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
            : "r"(a), "r"(b));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= r[i];
    if (x == 0x12345678u) out[0] = x;   // keep the chain live
}
__global__ void __launch_bounds__(256) k_bench_imad32(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
    uint32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= r[i];
    if (x == 0x12345678u) out[0] = x;
}
__global__ void __launch_bounds__(256) k_bench_fe(uint32_t* out, int iters, uint32_t seed, int square) {
    fe x, y;
#pragma unroll
    for (int i = 0; i < 8; i++) { x.v[i] = seed * (i + 1) + threadIdx.x; y.v[i] = seed * (i + 7) + blockIdx.x; }
    if (square) {
#pragma unroll 1
        for (int it = 0; it < iters; it++) { fe_sqr(x, x); fe_sqr(y, y); }
    } else {
#pragma unroll 1
        for (int it = 0; it < iters; it++) { fe_mul(x, x, y); fe_mul(y, y, x); }
    }
    uint32_t z = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) z ^= x.v[i] ^ y.v[i];
    if (z == 0x12345678u) out[0] = z;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side: context, workspace, pipeline
// ------------------------------------------------------------------------------------------------
#include "internal.h"

struct Precomp { const uint4* base; uint32_t stride; int c, W; };

#define LAUNCH_CHECK(ctx)            \
    do {                             \
        (ctx)->launches++;           \
        CK(ctx, cudaGetLastError()); \
    } while (0)

static int ensure(zk_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return ZK_OK;
    if (b.p) { CK(ctx, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t cap = bytes + bytes / 8 + 256;
    CK(ctx, cudaMalloc(&b.p, cap));
    b.cap = cap;
    return ZK_OK;
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }
static cudaError_t wait_main(zk_ctx* ctx);

// Every context owns four streams (main, two decode side streams, the high-priority one) and callers keep several
// contexts in flight; the driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8), and streams
// that share a queue serialise falsely (measured: 6 contexts in flight 2.67 ms/step with 8 queues, 1.52 with 32).  The
// variable is only read when the process creates its CUDA context, so set a default when the library is loaded; an
// explicit setting by the application wins.
__attribute__((constructor)) static void zk_default_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

extern "C" int zk_abi_version(void) { return ZK_ABI_VERSION; }

extern "C" const char* zk_status_str(int s) {
    switch (s) {
        case ZK_OK: return "ok";
        case ZK_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
        case ZK_ERR_INVALID_POINT: return "invalid ristretto255 encoding";
        case ZK_ERR_ARG: return "bad argument";
        case ZK_ERR_NOMEM: return "out of device memory";
        default: return "unknown status";
    }
}
extern "C" const char* zk_last_error(const zk_ctx* ctx) { return ctx ? ctx->err : ""; }

extern "C" int zk_ctx_create(int device, zk_ctx** out) {
    if (!out) return ZK_ERR_ARG;
    *out = nullptr;
    zk_ctx* ctx = new (std::nothrow) zk_ctx();
    if (!ctx) return ZK_ERR_NOMEM;
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 5 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev[i]);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_scan, cudaEventDisableTiming);
#if ZK_TAIL_HP
    if (e == cudaSuccess) {
        int lo = 0, hi = 0;
        e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->tail, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_acc, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_tail, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_pre, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_sort, cudaEventDisableTiming);
    }
#endif
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_out, 64);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        fprintf(stderr, "zkmsm: cannot create context on CUDA device %d: %s (there is no CPU fallback)\n", device,
                cudaGetErrorString(e));
        delete ctx;
        return ZK_ERR_CUDA;
    }
    *out = ctx;
    return ZK_OK;
}

extern "C" void zk_ctx_destroy(zk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2; i++) if (ctx->aux[i]) cudaStreamSynchronize(ctx->aux[i]);
    DevBuf* bufs[] = {&ctx->scalars, &ctx->comp, &ctx->dyn_table, &ctx->counts, &ctx->cursor, &ctx->offsets, &ctx->tiles,
                      &ctx->entries, &ctx->partials, &ctx->task_off, &ctx->tasks, &ctx->plan, &ctx->tree_a, &ctx->tree_w,
                      &ctx->out_ext, &ctx->out32, &ctx->bad, &ctx->seg, &ctx->batch_ext, &ctx->batch_out, &ctx->inv_scratch};
    for (DevBuf* b : bufs) if (b->p) cudaFree(b->p);
    for (int i = 0; i < 5; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    for (int i = 0; i < ZK_STAGE_SLOTS; i++) {
        if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    }
    for (int i = 0; i < 2; i++) { if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]); if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_scan) cudaEventDestroy(ctx->ev_scan);
    if (ctx->tail) { cudaStreamSynchronize(ctx->tail); cudaStreamDestroy(ctx->tail); }
    if (ctx->ev_acc) cudaEventDestroy(ctx->ev_acc);
    if (ctx->ev_tail) cudaEventDestroy(ctx->ev_tail);
    if (ctx->ev_pre) cudaEventDestroy(ctx->ev_pre);
    if (ctx->ev_sort) cudaEventDestroy(ctx->ev_sort);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int zk_ctx_sync(zk_ctx* ctx) {
    if (!ctx) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, wait_main(ctx));
    return ZK_OK;
}
extern "C" int zk_ctx_set_wait(zk_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 1) return ZK_ERR_ARG;
    ctx->wait_mode = mode;
    return ZK_OK;
}
extern "C" void* zk_ctx_stream(zk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int zk_ctx_set_window(zk_ctx* ctx, int c) {
    if (!ctx || (c != 0 && (c < 4 || c > 20))) return ZK_ERR_ARG;
    ctx->forced_window = c;
    return ZK_OK;
}
extern "C" int zk_ctx_set_profiling(zk_ctx* ctx, int on) {
    if (!ctx) return ZK_ERR_ARG;
    ctx->profiling = on;
#ifdef ZK_TIMELINE
    extern void zk_timeline_ref(zk_ctx*);
    if (on) zk_timeline_ref(ctx);
#endif
    return ZK_OK;
}
extern "C" int zk_ctx_last_phase_ms(zk_ctx* ctx, float out_ms[4]) {
    if (!ctx || !out_ms) return ZK_ERR_ARG;
    for (int i = 0; i < 4; i++) out_ms[i] = ctx->phase_ms[i];
    return ZK_OK;
}
extern "C" uint64_t zk_ctx_launch_count(const zk_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int zk_ctx_set_staging(zk_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return ZK_ERR_ARG;
    ctx->staging_mode = mode;
    return ZK_OK;
}
extern "C" uint64_t zk_ctx_staged_bytes(const zk_ctx* ctx) { return ctx ? ctx->staged_bytes : 0; }

// Page-lock a caller-owned host buffer so that uploads from it are true asynchronous DMA (what a caller with a
// long-lived Vec<u8> should do once, instead of paying the staging memcpy on every call).
extern "C" int zk_host_register(void* ptr, size_t bytes) {
    if (!ptr || !bytes) return ZK_ERR_ARG;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? ZK_ERR_NOMEM : ZK_ERR_CUDA; }
    return ZK_OK;
}
extern "C" int zk_host_unregister(void* ptr) {
    if (!ptr) return ZK_ERR_ARG;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return ZK_ERR_CUDA; }
    return ZK_OK;
}

// Wait for everything queued on the ctx's main stream: spinning in cudaStreamSynchronize (default, lowest latency), or
// parked on a blocking-sync event (zk_ctx_set_wait(ctx, 1)).  With several contexts in flight per GPU and several GPUs
// per host, spinning waiters occupy one core each and can starve the threads that feed the copies; a blocking wait
// costs ~0.3 ms of wake-up latency per call instead.
static cudaError_t wait_main(zk_ctx* ctx) {
    if (ctx->wait_mode == 0) return cudaStreamSynchronize(ctx->stream);
    if (!ctx->ev_done) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ctx->ev_done, ctx->stream);
    return e == cudaSuccess ? cudaEventSynchronize(ctx->ev_done) : e;
}

// ---- host -> device uploads -------------------------------------------------------------------------------------
static bool host_is_pageable(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// A few process-wide helper threads that share the staging memcpy()s with the calling thread: one core copies at
// ~10 GB/s, a x16 Gen5 link moves ~55 GB/s, so a single-threaded staging copy would be the bottleneck of an upload.
namespace {
struct CopyPool {
    static constexpr int MAX_HELPERS = 3;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    struct Job { uint8_t* dst; const uint8_t* src; size_t len; };
    Job jobs[MAX_HELPERS];
    int pending[MAX_HELPERS] = {0, 0, 0};      // 1 = posted, 2 = running
    uint64_t ticket = 0;
    std::thread th[MAX_HELPERS];
    int nhelpers = 0;
    bool started = false, quit = false;
    std::mutex user_mu;                         // one striped copy at a time (callers from several threads queue up)
    void start() {
        std::lock_guard<std::mutex> lk(mu);
        if (started) return;
        started = true;
        unsigned hw = std::thread::hardware_concurrency();
        nhelpers = hw >= 16 ? 3 : hw >= 8 ? 2 : hw >= 4 ? 1 : 0;
        for (int i = 0; i < nhelpers; i++) th[i] = std::thread([this, i] { run(i); });
    }
    void run(int i) {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return pending[i] == 1 || quit; });
            if (quit) return;
            Job j = jobs[i]; pending[i] = 2;
            lk.unlock();
            memcpy(j.dst, j.src, j.len);
            lk.lock();
            pending[i] = 0;
            cv_done.notify_all();
        }
    }
    void copy(uint8_t* dst, const uint8_t* src, size_t len) {
        if (!started) start();
        if (nhelpers == 0 || len < ((size_t)1 << 20) || !user_mu.try_lock()) { memcpy(dst, src, len); return; }
        const int parts = nhelpers + 1;
        const size_t per = ((len / parts) + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (int i = 0; i < nhelpers; i++) {
                size_t off = per * (i + 1);
                size_t l = off >= len ? 0 : (i == nhelpers - 1 ? len - off : (off + per > len ? len - off : per));
                jobs[i] = {dst + off, src + off, l};
                pending[i] = 1;
            }
            cv_work.notify_all();
        }
        memcpy(dst, src, per < len ? per : len);
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { for (int i = 0; i < nhelpers; i++) if (pending[i]) return false; return true; });
        }
        user_mu.unlock();
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(mu); quit = true; cv_work.notify_all(); }
        for (int i = 0; i < nhelpers; i++) if (th[i].joinable()) th[i].join();
    }
};
CopyPool g_copy_pool;
}  // namespace

// Asynchronous on `st` when the source is page-locked; from pageable memory the bytes go through the ctx's pinned
// ring (the calling thread copies chunk i+1 while chunk i is in flight), so `src` may be reused as soon as this returns.
static bool would_stage(const zk_ctx* ctx, const uint8_t* src, size_t bytes) {
    return ctx->staging_mode == 2 || (ctx->staging_mode == 0 && bytes >= ZK_STAGE_MIN_BYTES && host_is_pageable(src));
}
static int h2d(zk_ctx* ctx, void* dst, const uint8_t* src, size_t bytes, cudaStream_t st) {
    if (!bytes) return ZK_OK;
    if (!would_stage(ctx, src, bytes)) {
        CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return ZK_OK;
    }
    for (size_t off = 0; off < bytes; off += ZK_STAGE_SLOT_BYTES) {
        const size_t len = bytes - off < ZK_STAGE_SLOT_BYTES ? bytes - off : ZK_STAGE_SLOT_BYTES;
        const unsigned slot = ctx->stage_next++ % ZK_STAGE_SLOTS;
        if (!ctx->stage[slot]) {
            CK(ctx, cudaMallocHost((void**)&ctx->stage[slot], ZK_STAGE_SLOT_BYTES));
            CK(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[slot], cudaEventDisableTiming));
        }
        if (ctx->stage_busy[slot]) CK(ctx, cudaEventSynchronize(ctx->stage_ev[slot]));
        g_copy_pool.copy(ctx->stage[slot], src + off, len);
        CK(ctx, cudaMemcpyAsync((uint8_t*)dst + off, ctx->stage[slot], len, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaEventRecord(ctx->stage_ev[slot], st));
        ctx->stage_busy[slot] = true;
        ctx->staged_bytes += len;
    }
    return ZK_OK;
}

// After a failure between the fork onto the side streams and the join: let everything queued drain, so that no copy
// still reads the caller's buffers and no decoder still writes ctx->bad when the next call starts.
static void quiesce(zk_ctx* ctx) {
    if (ctx->tail) cudaStreamSynchronize(ctx->tail);     // a failed call may have left the hop back to the main stream unqueued
    for (int i = 0; i < 2; i++) if (ctx->aux[i]) cudaStreamSynchronize(ctx->aux[i]);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    ctx->join_aux = false;
    cudaGetLastError();
}

// Window width by size, measured on B200 (tools/sweep_windows.py, profiles/README.md).  Widths are balanced across
// windows (window_geom), so the curve is smooth in c; small MSMs sit on the serial tail and only care about
// keeping the tree shallow.
extern "C" int zk_pick_window(size_t n) {
    if (n < ((size_t)1 << 11)) return 10;
    if (n < ((size_t)1 << 15)) return 13;
    if (n < ((size_t)1 << 20)) return 15;
    if (n < ((size_t)1 << 23)) return 16;
    if (n < ((size_t)1 << 24)) return 18;     // block-scale sizes (tools/big_sweep.py)
    if (n < ((size_t)1 << 26)) return 19;
    return 20;
}

// Batches are throughput-bound (many MSMs hide each other's serial tails), so the width minimises the multiply
// count W * (7 n + 36 * 2^(c-1)): n mixed adds per window plus two quad additions per bucket in the tree.
static double window_cost(double n, int c, bool shared_buckets) {
    double W = windows_for_width(c), B = (double)(1u << (max_width((int)W) - 1));
    return W * 7.0 * n + (shared_buckets ? 1.0 : W) * 36.0 * B;
}
static int pick_window_batch(size_t n_avg) {
    int best = 4; double best_cost = 1e300;
    for (int c = 4; c <= 16; c++) {
        double cost = window_cost((double)(n_avg ? n_avg : 1), c, false);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

// ---- tables ----
// Thread-safety: a zk_table may be READ by several contexts/threads at once (MSMs over it), but appends, clear,
// precompute and destroy must not run concurrently with any other use of the same table (they may move its storage).
extern "C" int zk_table_create(zk_ctx* ctx, size_t capacity, zk_table** out) {
    if (!ctx || !out) return ZK_ERR_ARG;
    *out = nullptr;
    zk_table* t = new (std::nothrow) zk_table();
    if (!t) return ZK_ERR_NOMEM;
    t->device = ctx->device;
    t->cap = capacity ? capacity : 1;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&t->d, t->cap * 96);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "zk_table_create: %s", cudaGetErrorString(e));
        delete t;
        return e == cudaErrorMemoryAllocation ? ZK_ERR_NOMEM : ZK_ERR_CUDA;
    }
    *out = t;
    return ZK_OK;
}
extern "C" void zk_table_destroy(zk_table* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaDeviceSynchronize();                 // other contexts may still have MSMs over this table in flight
    if (t->pre) cudaFree(t->pre);
    if (t->d) cudaFree(t->d);
    delete t;
}
extern "C" size_t zk_table_len(const zk_table* t) { return t ? t->len : 0; }
extern "C" size_t zk_table_capacity(const zk_table* t) { return t ? t->cap : 0; }
static void table_drop_precomp(zk_table* t) {
    if (t->pre) { cudaSetDevice(t->device); cudaDeviceSynchronize(); cudaFree(t->pre); }
    t->pre = nullptr; t->pre_len = 0; t->pre_c = t->pre_W = 0;
}
extern "C" void zk_table_clear(zk_table* t) { if (t) { t->len = 0; table_drop_precomp(t); } }
extern "C" int zk_table_precomputed_window(const zk_table* t) { return t && t->pre ? t->pre_c : 0; }

// Room for `need` rows.  Growth moves the storage: work queued by OTHER contexts over the old buffer is drained first
// (cudaDeviceSynchronize), so a table shared between contexts stays valid for everything already submitted.
static int table_reserve(zk_ctx* ctx, zk_table* t, size_t need) {
    if (need <= t->cap) return ZK_OK;
    size_t cap = t->cap * 2 > need ? t->cap * 2 : need;
    uint4* nd = nullptr;
    CK(ctx, cudaMalloc((void**)&nd, cap * 96));
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess && t->len) e = cudaMemcpyAsync(nd, t->d, t->len * 96, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(nd); CK(ctx, e); }
    CK(ctx, cudaFree(t->d));
    t->d = nd; t->cap = cap;
    return ZK_OK;
}
// `added` rows were appended and validated: only now does an existing window expansion become stale.
static void table_commit(zk_table* t, size_t added) {
    if (added) { table_drop_precomp(t); t->len += added; }
}

static int read_bad_index(zk_ctx* ctx, size_t* bad_index) {
    CK(ctx, cudaMemcpyAsync(ctx->h_out + 32, ctx->bad.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long b; memcpy(&b, ctx->h_out + 32, 8);
    if (b != ~0ull) { if (bad_index) *bad_index = (size_t)b; return ZK_ERR_INVALID_POINT; }
    return ZK_OK;
}

// Decoder launches are sized in whole waves of the device (4 resident 128-thread CTAs per SM at 106 registers), so
// that only the last chunk of an upload has a partial tail wave.
static size_t decode_chunk(const zk_ctx* ctx, size_t n) {
    if (n <= ((size_t)1 << 16)) return n;
    const size_t wave = (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 4 * 128;
    size_t waves = (n / 8 + wave / 2) / wave;
    if (waves < 1) waves = 1;
    return waves * wave;
}

// decompress n encodings at `src_dev` into table rows [dst_row, dst_row+n); returns INVALID_POINT + index on reject.
static int decompress_into(zk_ctx* ctx, const void* src_dev, size_t n, uint4* table, size_t dst_row, size_t* bad_index) {
    if (n == 0) return ZK_OK;
    TRY(ensure(ctx, ctx->bad, 8));
    CK(ctx, cudaMemsetAsync(ctx->bad.p, 0xff, 8, ctx->stream));
    k_decompress<<<grid_for(ZK_DEC_PAIR ? (n + 1) / 2 : n, 128), 128, 0, ctx->stream>>>((const uint4*)src_dev, n, table + dst_row * 6,
                                                             (unsigned long long*)ctx->bad.p, 0ull);
    LAUNCH_CHECK(ctx);
    return read_bad_index(ctx, bad_index);
}

extern "C" int zk_table_append_compressed_dev(zk_ctx* ctx, zk_table* t, const void* points32_dev, size_t n, size_t* bad_index) {
    if (!ctx || !t || (!points32_dev && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(table_reserve(ctx, t, t->len + n));
    TRY(decompress_into(ctx, points32_dev, n, t->d, t->len, bad_index));
    table_commit(t, n);
    return ZK_OK;
}
extern "C" int zk_table_append_compressed(zk_ctx* ctx, zk_table* t, const uint8_t* points32_host, size_t n, size_t* bad_index) {
    if (!ctx || !t || (!points32_host && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->comp, n * 32));
    TRY(h2d(ctx, ctx->comp.p, points32_host, n * 32, ctx->stream));
    return zk_table_append_compressed_dev(ctx, t, ctx->comp.p, n, bad_index);
}
// ext (n x 128 B in HBM) -> rows [dst_row, dst_row + n) with the batched inversion; `bad` may be null
static int normalise_into(zk_ctx* ctx, const void* ext_dev, size_t n, uint4* table, size_t dst_row, unsigned long long* bad) {
    k_ext_to_niels<<<grid_for((n + INV_BATCH - 1) / INV_BATCH, 128), 128, 0, ctx->stream>>>((const uint4*)ext_dev, n, table + dst_row * 6, bad);
    LAUNCH_CHECK(ctx);
    return ZK_OK;
}
extern "C" int zk_table_append_uniform_dev(zk_ctx* ctx, zk_table* t, const void* bytes64_dev, size_t n) {
    if (!ctx || !t || (!bytes64_dev && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(table_reserve(ctx, t, t->len + n));
    if (n) {
        TRY(ensure(ctx, ctx->inv_scratch, n * 128));
        k_map_uniform<<<grid_for(n, 128), 128, 0, ctx->stream>>>((const uint4*)bytes64_dev, n, (uint4*)ctx->inv_scratch.p);
        LAUNCH_CHECK(ctx);
        TRY(normalise_into(ctx, ctx->inv_scratch.p, n, t->d, t->len, nullptr));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    table_commit(t, n);
    return ZK_OK;
}
extern "C" int zk_table_append_uniform(zk_ctx* ctx, zk_table* t, const uint8_t* bytes64_host, size_t n) {
    if (!ctx || !t || (!bytes64_host && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->comp, n * 64));
    TRY(h2d(ctx, ctx->comp.p, bytes64_host, n * 64, ctx->stream));
    return zk_table_append_uniform_dev(ctx, t, ctx->comp.p, n);
}
extern "C" int zk_table_precompute(zk_ctx* ctx, zk_table* t, int c) {
    if (!ctx || !t || (c != 0 && (c < 4 || c > 20))) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    table_drop_precomp(t);
    const size_t len = t->len;
    if (len == 0) return ZK_OK;
    if (c == 0) {              // measured optimum (tools/perf_precomp.py, perf_batch_pre.py): one bit more than log2(len)
        int lg = 0; while (((size_t)2 << lg) <= len) lg++;
        c = lg + 1; if (c < 8) c = 8; if (c > 20) c = 20;
    }
    const int W = windows_for_width(c);
    c = max_width(W);
    if ((unsigned long long)len * W >= (1ull << 31)) return ZK_ERR_ARG;
    uint4* pre = nullptr;
    TRY(ensure(ctx, ctx->inv_scratch, len * (size_t)(W - 1) * 128));      // the doubling chains' extended points, kept in the workspace
    uint4* scratch = (uint4*)ctx->inv_scratch.p;
    CK(ctx, cudaMalloc((void**)&pre, len * W * 96));
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(pre, t->d, len * 96, cudaMemcpyDeviceToDevice, st);             // window 0 = the points themselves
    if (e == cudaSuccess) {
        const size_t m = len * (size_t)(W - 1);
        k_precomp_double<<<grid_for(len, 128), 128, 0, st>>>(t->d, len, W, scratch);
        k_ext_to_niels<<<grid_for((m + INV_BATCH - 1) / INV_BATCH, 128), 128, 0, st>>>(scratch, m, pre + len * 6, nullptr);
        ctx->launches += 2;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(pre); CK(ctx, e); }
    t->pre = pre; t->pre_len = len; t->pre_c = c; t->pre_W = W;
    return ZK_OK;
}

static int append_extended_common(zk_ctx* ctx, zk_table* t, const void* ext128_dev, size_t n, size_t* bad_index, bool checked) {
    if (!ctx || !t || (!ext128_dev && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(table_reserve(ctx, t, t->len + n));
    if (n == 0) return ZK_OK;
    TRY(ensure(ctx, ctx->bad, 8));
    CK(ctx, cudaMemsetAsync(ctx->bad.p, 0xff, 8, ctx->stream));
    if (checked) {
        k_from_extended<<<grid_for(n, 128), 128, 0, ctx->stream>>>((const uint4*)ext128_dev, n, t->d + t->len * 6,
                                                                    (unsigned long long*)ctx->bad.p);
        LAUNCH_CHECK(ctx);
    } else {
        TRY(normalise_into(ctx, ext128_dev, n, t->d, t->len, (unsigned long long*)ctx->bad.p));
    }
    TRY(read_bad_index(ctx, bad_index));
    table_commit(t, n);
    return ZK_OK;
}
extern "C" int zk_table_append_extended_dev(zk_ctx* ctx, zk_table* t, const void* ext128_dev, size_t n, size_t* bad_index) {
    return append_extended_common(ctx, t, ext128_dev, n, bad_index, true);
}
extern "C" int zk_table_append_extended_unchecked_dev(zk_ctx* ctx, zk_table* t, const void* ext128_dev, size_t n, size_t* bad_index) {
    return append_extended_common(ctx, t, ext128_dev, n, bad_index, false);
}
static int append_extended_host(zk_ctx* ctx, zk_table* t, const uint8_t* ext128_host, size_t n, size_t* bad_index, bool checked) {
    if (!ctx || !t || (!ext128_host && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->comp, n * 128));
    TRY(h2d(ctx, ctx->comp.p, ext128_host, n * 128, ctx->stream));
    return append_extended_common(ctx, t, ctx->comp.p, n, bad_index, checked);
}
extern "C" int zk_table_append_extended(zk_ctx* ctx, zk_table* t, const uint8_t* ext128_host, size_t n, size_t* bad_index) {
    return append_extended_host(ctx, t, ext128_host, n, bad_index, true);
}
extern "C" int zk_table_append_extended_unchecked(zk_ctx* ctx, zk_table* t, const uint8_t* ext128_host, size_t n, size_t* bad_index) {
    return append_extended_host(ctx, t, ext128_host, n, bad_index, false);
}
extern "C" int zk_table_compress_dev(zk_ctx* ctx, const zk_table* t, size_t offset, size_t n, void* out32_dev) {
    if (!ctx || !t || (!out32_dev && n) || offset > t->len || n > t->len - offset) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    if (n) {
        k_compress_table<<<grid_for(n, 128), 128, 0, ctx->stream>>>(t->d + offset * 6, n, (uint4*)out32_dev);
        LAUNCH_CHECK(ctx);
    }
    return ZK_OK;
}
extern "C" int zk_table_compress(zk_ctx* ctx, const zk_table* t, size_t offset, size_t n, uint8_t* out32_host) {
    if (!ctx || !t || (!out32_host && n)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->comp, n * 32 + 32));
    TRY(zk_table_compress_dev(ctx, t, offset, n, ctx->comp.p));
    CK(ctx, cudaMemcpyAsync(out32_host, ctx->comp.p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// ---- the MSM pipeline (asynchronous on ctx->stream) ----
// A plan fixes the geometry of one pass and makes sure the workspace exists.  It is computed BEFORE anything is
// queued on the side streams, so that an argument or out-of-memory failure never leaves copies or decoders in flight.
struct MsmPlan {
    size_t n = 0, nmsm = 1;
    int W1 = 0, c = 0, WB = 0;           // digit windows per scalar, widest window, bucket-space windows per MSM
    size_t W = 0, B = 0, NB = 0, ntiles = 0, max_tasks = 0, m1 = 0;
    bool use_pc = false; Precomp pc{};
};

// Does the window expansion pay for an n-term MSM?  Its width was chosen for the whole table; a short slice of a long
// table would sweep 2^(pre_c - 1) buckets for a handful of entries, so compare the multiply counts of both routes.
static bool precomp_pays(size_t n, int pre_c) {
    return window_cost((double)n, pre_c, true) <= window_cost((double)n, zk_pick_window(n), false);
}
static bool table_precomp(const zk_table* t, size_t offset, size_t n, Precomp* pc) {
    if (!t->pre || offset + n > t->pre_len || !precomp_pays(n, t->pre_c)) return false;
    pc->base = t->pre + offset * 6; pc->stride = (uint32_t)t->pre_len; pc->c = t->pre_c; pc->W = t->pre_W;
    return true;
}

static int msm_plan(zk_ctx* ctx, size_t n, size_t nmsm, const Precomp* pc, MsmPlan* p) {
    p->n = n; p->nmsm = nmsm; p->use_pc = pc != nullptr;
    if (pc) p->pc = *pc;
    if (n == 0) return ZK_OK;
    if (n >= (1ull << 31)) return ZK_ERR_ARG;
    // precomputed-window table: its width is fixed, every window of an MSM lands in one shared bucket set, no Horner
    const int c_req = pc ? pc->c : ctx->forced_window ? ctx->forced_window : (nmsm > 1 ? pick_window_batch(n / nmsm) : zk_pick_window(n));
    p->W1 = pc ? pc->W : windows_for_width(c_req);   // digit windows per scalar (balanced widths, see window_geom)
    p->c = max_width(p->W1);                           // widest window -> buckets per window
    p->WB = pc ? 1 : p->W1;
    p->W = (size_t)p->WB * nmsm;                       // bucket-space windows in the whole batch
    p->B = (size_t)1 << (p->c - 1);
    p->NB = p->B * p->W;
    if (p->NB >= (1ull << 31)) return ZK_ERR_ARG;
    p->ntiles = (p->NB + SCAN_TILE - 1) / SCAN_TILE;
    if ((unsigned long long)n * p->W1 >= (1ull << 32)) return ZK_ERR_ARG;   // entry positions are 32-bit: shard larger MSMs
    p->max_tasks = p->NB + (n * (size_t)p->W1) / TASK_LEN;
    p->m1 = (p->B + REDUCE_RADIX - 1) / REDUCE_RADIX;
    TRY(ensure(ctx, ctx->counts, p->NB * 4));
    TRY(ensure(ctx, ctx->cursor, p->NB * 4));
    TRY(ensure(ctx, ctx->offsets, (p->NB + 1) * 4));
    TRY(ensure(ctx, ctx->task_off, (p->NB + 1) * 4));
    TRY(ensure(ctx, ctx->tiles, p->ntiles * 8));
    TRY(ensure(ctx, ctx->plan, (TASK_LEN + 2) * 4));
    TRY(ensure(ctx, ctx->tasks, p->max_tasks * 8));
    TRY(ensure(ctx, ctx->entries, n * p->W1 * 4));
    TRY(ensure(ctx, ctx->partials, p->max_tasks * 128));
    TRY(ensure(ctx, ctx->tree_a, 2 * p->m1 * p->W * 128));     // ping-pong halves
    TRY(ensure(ctx, ctx->tree_w, 2 * p->m1 * p->W * 128));
    return ZK_OK;
}

// scalars: n*32 B in HBM.  Point i lives at tab_a[i] for i < split, tab_b[i - split] otherwise.
// Batch mode (nmsm > 1): seg_dev = nmsm+1 offsets in HBM; out_ext_dev receives nmsm extended points.
// fused_out32 != nullptr (single MSM only): the tail also encodes, into that device buffer (k_combine_encode).
// Dynamic (compressed) points whose upload + decode is queued INSIDE the pipeline, after the scan, with their share of
// the digit scatter fused into the decoder (k_decompress_scatter).  Worth it from ZK_FUSE_SCATTER_MIN points on: below,
// making the decoder wait for the histogram costs more latency than the hidden scatter saves.
#ifndef ZK_FUSE_SCATTER
#define ZK_FUSE_SCATTER 1
#endif
constexpr size_t ZK_FUSE_SCATTER_MIN = (size_t)1 << 17;
struct FusedDyn { const uint8_t* points32_host = nullptr; size_t n = 0; };
static inline bool fuse_scatter_pays(size_t n_dyn) { return ZK_FUSE_SCATTER && !ZK_DEC_PAIR && n_dyn >= ZK_FUSE_SCATTER_MIN; }

#if !ZK_DEC_PAIR
// Chunks alternate between the two side streams (copy of chunk i+1 under the decode of chunk i); each chunk's kernel waits
// for `ready` (histogram + scan done: cursor holds the bucket offsets).  Term k of the dynamic part is term n_static + k
// of the MSM.  ctx->ev_fork must have been recorded on the main stream after ctx->bad was reset.
static int launch_decode_scatter(zk_ctx* ctx, const FusedDyn& fd, size_t n_static, const void* scalars_dev, int c, int W1, cudaEvent_t ready) {
    const size_t chunk = decode_chunk(ctx, fd.n);
    int which = 0;
    ctx->join_aux = true;          // from here on an error return must quiesce() the side streams
    for (size_t lo = 0; lo < fd.n; lo += chunk, which ^= 1) {
        size_t cnt = fd.n - lo < chunk ? fd.n - lo : chunk;
        cudaStream_t sa = ctx->aux[which];
        if (lo < 2 * chunk) CK(ctx, cudaStreamWaitEvent(sa, ctx->ev_fork, 0));
        TRY(h2d(ctx, (uint8_t*)ctx->comp.p + lo * 32, fd.points32_host + lo * 32, cnt * 32, sa));
        if (lo < 2 * chunk) CK(ctx, cudaStreamWaitEvent(sa, ready, 0));
        k_decompress_scatter<<<grid_for(cnt, 128), 128, 0, sa>>>((const uint4*)ctx->comp.p + lo * 2, cnt, (uint4*)ctx->dyn_table.p + lo * 6,
                                                                (unsigned long long*)ctx->bad.p, (unsigned long long)lo,
                                                                (const uint4*)scalars_dev + (n_static + lo) * 2, (uint32_t)(n_static + lo), c, W1,
                                                                (uint32_t*)ctx->cursor.p, (uint32_t*)ctx->entries.p);
        LAUNCH_CHECK(ctx);
    }
    for (int i = 0; i < 2; i++) CK(ctx, cudaEventRecord(ctx->ev_join[i], ctx->aux[i]));
    return ZK_OK;
}
#endif

static int msm_enqueue(zk_ctx* ctx, const MsmPlan& p, const void* scalars_dev, const uint4* tab_a, const uint4* tab_b, size_t split,
                       void* out_ext_dev, const uint32_t* seg_dev = nullptr, bool shared_points = false, void* fused_out32 = nullptr,
                       const FusedDyn* fd = nullptr) {
    cudaStream_t st = ctx->stream;
    const size_t n = p.n, nmsm = p.nmsm;
    if (n == 0) {
        if (ctx->join_aux) {
            for (int i = 0; i < 2; i++) CK(ctx, cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
            ctx->join_aux = false;
        }
        k_set_identity_batch<<<grid_for(nmsm, 128), 128, 0, st>>>((uint4*)out_ext_dev, nmsm);
        LAUNCH_CHECK(ctx);
        if (fused_out32) {
            k_ext_sum_encode<<<1, 32, 0, st>>>((const uint4*)out_ext_dev, 1, (uint4*)fused_out32);
            LAUNCH_CHECK(ctx);
        }
        return ZK_OK;
    }
    const int W1 = p.W1, c = p.c, WB = p.WB;
    const size_t W = p.W, B = p.B, NB = p.NB, ntiles = p.ntiles, max_tasks = p.max_tasks, m1 = p.m1;
    const uint32_t win_stride = p.use_pc ? p.pc.stride : 0u;
    if (p.use_pc) { tab_a = p.pc.base; tab_b = p.pc.base; split = (size_t)0xffffffffu; }

    if (ctx->profiling) CK(ctx, cudaEventRecord(ctx->ev[1], st));
#if ZK_TAIL_HP >= 2
    // the digit sort also runs on the high-priority stream: only the bulk kernels (accumulate, decode) stay at low priority
    cudaStream_t st_bulk = st;
    CK(ctx, cudaEventRecord(ctx->ev_pre, st));
    CK(ctx, cudaStreamWaitEvent(ctx->tail, ctx->ev_pre, 0));
    st = ctx->tail;
#endif
    CK(ctx, cudaMemsetAsync(ctx->counts.p, 0, NB * 4, st));
    CK(ctx, cudaMemsetAsync(ctx->plan.p, 0, (TASK_LEN + 2) * 4, st));
    k_digit_hist<<<grid_for(n, ZK_SORT_BLOCK), ZK_SORT_BLOCK, 0, st>>>((const uint4*)scalars_dev, n, c, W1, seg_dev, (uint32_t)nmsm, win_stride, (uint32_t*)ctx->counts.p);
    LAUNCH_CHECK(ctx);
    k_scan_tile_sums<<<(unsigned)ntiles, SCAN_BLOCK, 0, st>>>((const uint32_t*)ctx->counts.p, NB, (unsigned long long*)ctx->tiles.p,
                                                        (uint32_t*)ctx->plan.p);
    LAUNCH_CHECK(ctx);
    k_scan_tiles<<<1, SCAN_BLOCK, 0, st>>>((unsigned long long*)ctx->tiles.p, ntiles, (uint32_t*)ctx->plan.p);
    LAUNCH_CHECK(ctx);
    k_scan_apply<<<(unsigned)ntiles, SCAN_BLOCK, 0, st>>>((const uint32_t*)ctx->counts.p, NB, (const unsigned long long*)ctx->tiles.p,
                                                    (uint32_t*)ctx->offsets.p, (uint32_t*)ctx->cursor.p, (uint32_t*)ctx->task_off.p,
                                                    (uint32_t*)ctx->plan.p, (uint2*)ctx->tasks.p);
    LAUNCH_CHECK(ctx);
    // terms whose points are cached are scattered here; dynamic terms with a fused decoder (fd) scatter themselves
    const size_t n_scatter = fd ? n - fd->n : n;
    if (n_scatter) {
        k_digit_scatter<<<grid_for(n_scatter, ZK_SORT_BLOCK), ZK_SORT_BLOCK, 0, st>>>((const uint4*)scalars_dev, n_scatter, c, W1, seg_dev, (uint32_t)nmsm,
                                                           shared_points ? 1 : 0, win_stride, (uint32_t*)ctx->cursor.p, (uint32_t*)ctx->entries.p);
        LAUNCH_CHECK(ctx);
    }
    if (ctx->profiling) CK(ctx, cudaEventRecord(ctx->ev[2], st));
#if !ZK_DEC_PAIR
    if (fd) {
        CK(ctx, cudaEventRecord(ctx->ev_scan, st));
        TRY(launch_decode_scatter(ctx, *fd, n - fd->n, scalars_dev, c, W1, ctx->ev_scan));
    }
#endif
#if ZK_TAIL_HP >= 2
    CK(ctx, cudaEventRecord(ctx->ev_sort, st));
    st = st_bulk;
    CK(ctx, cudaStreamWaitEvent(st, ctx->ev_sort, 0));
#endif

    if (ctx->join_aux) {          // the point table is being produced on the side streams
        for (int i = 0; i < 2; i++) CK(ctx, cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
        ctx->join_aux = false;
    }
    // grid sized for the worst case; threads past the device-side task count exit at once
    k_bucket_accum<<<grid_for(max_tasks, 128), 128, 0, st>>>(tab_a, tab_b, (uint32_t)split, (const uint32_t*)ctx->entries.p,
                                                              (const uint32_t*)ctx->offsets.p, (const uint32_t*)ctx->task_off.p,
                                                              (const uint2*)ctx->tasks.p, (const uint32_t*)ctx->plan.p,
                                                              (uint4*)ctx->partials.p);
    LAUNCH_CHECK(ctx);
    if (ctx->profiling) CK(ctx, cudaEventRecord(ctx->ev[3], st));

    // the phases after the accumulation run on the context's high-priority stream (when it has one), ordered after
    // the accumulation by an event, and the main stream is made to wait for them: callers still see one stream
    cudaStream_t st_main = st;
    if (ctx->tail) {
        CK(ctx, cudaEventRecord(ctx->ev_acc, st));
        CK(ctx, cudaStreamWaitEvent(ctx->tail, ctx->ev_acc, 0));
        st = ctx->tail;
    }
    // radix-8 tree, per window, until one node is left
    const uint4* a_in = (const uint4*)ctx->partials.p;
    const uint4* w_in = nullptr;
    const uint32_t* toff = (const uint32_t*)ctx->task_off.p;
    size_t m_in = B; int log2_wc = 0; int half = 0;
    while (true) {
        size_t m_out = (m_in + REDUCE_RADIX - 1) / REDUCE_RADIX;
        uint4* a_out = (uint4*)ctx->tree_a.p + (size_t)half * m1 * W * 8;
        uint4* w_out = (uint4*)ctx->tree_w.p + (size_t)half * m1 * W * 8;
        if (toff == nullptr && m_out * W <= TREE_WARP_MAX_NODES)      // few nodes: latency-bound, spend lanes on depth
            k_tree_level_warp<<<grid_for(m_out * W * 32, ZK_TREE_BLOCK), ZK_TREE_BLOCK, 0, st>>>(a_in, w_in, m_in, m_out, (int)W, log2_wc, a_out, w_out);
        else
            k_tree_level_quad<<<grid_for(m_out * W * 4, ZK_TREE_BLOCK), ZK_TREE_BLOCK, 0, st>>>(a_in, w_in, toff, m_in, m_out, (int)W, log2_wc, a_out, w_out);
        LAUNCH_CHECK(ctx);
        a_in = a_out; w_in = w_out; toff = nullptr; m_in = m_out; log2_wc += REDUCE_RADIX_LOG2; half ^= 1;
        if (m_out == 1) break;
    }
    if (fused_out32 && nmsm == 1)
        k_combine_encode<<<1, 32, 0, st>>>(w_in, WB, W1, (uint4*)out_ext_dev, (uint4*)fused_out32);
    else
        k_window_combine<<<grid_for(nmsm, 8), 32, 0, st>>>(w_in, WB, W1, (uint32_t)nmsm, (uint4*)out_ext_dev);
    LAUNCH_CHECK(ctx);
    if (ctx->tail) {
        CK(ctx, cudaEventRecord(ctx->ev_tail, st));
        CK(ctx, cudaStreamWaitEvent(st_main, ctx->ev_tail, 0));
    }
    return ZK_OK;
}

// Read the 32-byte encoding in ctx->out32 back (after encoding the sum of g partials first, unless the pipeline's fused
// tail already did: ext_dev == nullptr).
static int finish_encode(zk_ctx* ctx, const void* ext_dev, size_t g, uint8_t out32[32]) {
    TRY(ensure(ctx, ctx->out32, 32));
    if (ext_dev) {
        // one warp of latency-bound work: on the high-priority stream, so that it does not wait for a free CTA slot
        // behind other contexts' bulk grids
        cudaStream_t s = ctx->stream;
        if (ctx->tail) {
            CK(ctx, cudaEventRecord(ctx->ev_pre, ctx->stream));
            CK(ctx, cudaStreamWaitEvent(ctx->tail, ctx->ev_pre, 0));
            s = ctx->tail;
        }
        k_ext_sum_encode<<<1, 32, 0, s>>>((const uint4*)ext_dev, g, (uint4*)ctx->out32.p);
        LAUNCH_CHECK(ctx);
        if (ctx->tail) {
            CK(ctx, cudaEventRecord(ctx->ev_tail, s));
            CK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_tail, 0));
        }
    }
    if (ctx->profiling) CK(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    CK(ctx, cudaMemcpyAsync(ctx->h_out, ctx->out32.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, wait_main(ctx));
    memcpy(out32, ctx->h_out, 32);
    if (ctx->profiling) {
        for (int i = 0; i < 4; i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) ctx->phase_ms[i] = ms;
        }
    }
    return ZK_OK;
}

extern "C" int zk_msm_table_dev(zk_ctx* ctx, const void* scalars32_dev, const zk_table* t, size_t offset, size_t n, void* out_ext128_dev) {
    if (!ctx || !t || !out_ext128_dev || (!scalars32_dev && n) || offset > t->len || n > t->len - offset) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    Precomp pc;
    const bool use_pc = table_precomp(t, offset, n, &pc);
    MsmPlan plan;
    TRY(msm_plan(ctx, n, 1, use_pc ? &pc : nullptr, &plan));
    if (ctx->profiling) { CK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream)); }
    int rc = msm_enqueue(ctx, plan, scalars32_dev, t->d + offset * 6, nullptr, n, out_ext128_dev);
    if (rc != ZK_OK) quiesce(ctx);      // the phases hop between two streams: leave none of them running behind an error
    return rc;
}

#ifdef ZK_TIMELINE
// Developer build only (-DZK_TIMELINE): with profiling on, every zk_msm_table_dev + zk_ext_sum_compress_dev pair prints
// its phase boundaries (call, sort start, sort end, accumulate end, tree + tail end) in ms since a process-wide
// reference event -- a timeline of several contexts in flight without a tracing tool.
static cudaEvent_t g_tl_ref; static bool g_tl_set = false;
#endif
#ifdef ZK_TIMELINE
void zk_timeline_ref(zk_ctx* ctx) {
    if (g_tl_set) return;
    cudaEventCreate(&g_tl_ref); cudaEventRecord(g_tl_ref, ctx->stream); cudaEventSynchronize(g_tl_ref); g_tl_set = true;
}
#endif
extern "C" int zk_ext_sum_compress_dev(zk_ctx* ctx, const void* ext128_dev, size_t g, uint8_t out32[32]) {
    if (!ctx || (!ext128_dev && g) || !out32) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    int prof = ctx->profiling; ctx->profiling = 0;
#ifdef ZK_TIMELINE
    if (prof) cudaEventRecord(ctx->ev[4], ctx->stream);
#endif
    int rc = finish_encode(ctx, ext128_dev, g, out32);
    ctx->profiling = prof;
#ifdef ZK_TIMELINE
    if (prof && g_tl_set) {
        float t[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; i++) cudaEventElapsedTime(&t[i], g_tl_ref, ctx->ev[i]);
        fprintf(stderr, "TL %p %.4f %.4f %.4f %.4f %.4f\n", (void*)ctx, t[0], t[1], t[2], t[3], t[4]);
    }
#endif
    return rc;
}

// Upload n compressed points from the host and decode them into ctx->dyn_table, in chunks alternating between the
// two side streams so that the copy of chunk i+1 overlaps the decode of chunk i.  The main stream is free to upload
// scalars and sort digits meanwhile; msm_enqueue() joins the side streams right before the accumulation.
// ctx->bad (lowest rejected index) must have been reset on the main stream.  Batch mode: seg/M/bad_msm mark the MSM.
static int start_upload_decode(zk_ctx* ctx, const uint8_t* points32_host, size_t n, const uint32_t* seg_dev, uint32_t M,
                               uint32_t* bad_msm_dev) {
    if (n == 0) return ZK_OK;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaEventRecord(ctx->ev_fork, st));
    const size_t chunk = decode_chunk(ctx, n);
    int which = 0;
    ctx->join_aux = true;          // from here on an error return must quiesce() the side streams
    for (size_t lo = 0; lo < n; lo += chunk, which ^= 1) {
        size_t cnt = n - lo < chunk ? n - lo : chunk;
        cudaStream_t sa = ctx->aux[which];
        if (lo < 2 * chunk) CK(ctx, cudaStreamWaitEvent(sa, ctx->ev_fork, 0));
        TRY(h2d(ctx, (uint8_t*)ctx->comp.p + lo * 32, points32_host + lo * 32, cnt * 32, sa));
        k_decompress<<<grid_for(ZK_DEC_PAIR ? (cnt + 1) / 2 : cnt, 128), 128, 0, sa>>>((const uint4*)ctx->comp.p + lo * 2, cnt, (uint4*)ctx->dyn_table.p + lo * 6,
                                                        (unsigned long long*)ctx->bad.p, (unsigned long long)lo, seg_dev, M, bad_msm_dev);
        LAUNCH_CHECK(ctx);
    }
    for (int i = 0; i < 2; i++) CK(ctx, cudaEventRecord(ctx->ev_join[i], ctx->aux[i]));
    return ZK_OK;
}

// Everything of a mixed (static table prefix + dynamic compressed suffix) MSM up to the extended result in
// ctx->out_ext and the reject index in ctx->h_out[32..40): queued, not waited for.
static int enqueue_partial_impl(zk_ctx* ctx, const zk_host_piece* pieces, int npieces, const zk_table* t, size_t offset, size_t n_static,
                                const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host, size_t n_dyn, bool fuse_encode) {
    const size_t n = n_static + n_dyn;
    Precomp pc;
    const bool use_pc = n_dyn == 0 && n_static && table_precomp(t, offset, n_static, &pc);
    MsmPlan plan;
    TRY(msm_plan(ctx, n, 1, use_pc ? &pc : nullptr, &plan));
    TRY(ensure(ctx, ctx->scalars, n * 32));
    TRY(ensure(ctx, ctx->out_ext, 128));
    TRY(ensure(ctx, ctx->out32, 32));
    TRY(ensure(ctx, ctx->comp, n_dyn * 32));
    TRY(ensure(ctx, ctx->dyn_table, n_dyn * 96));
    TRY(ensure(ctx, ctx->bad, 8));
    cudaStream_t st = ctx->stream;
    if (ctx->profiling) CK(ctx, cudaEventRecord(ctx->ev[0], st));
    CK(ctx, cudaMemsetAsync(ctx->bad.p, 0xff, 8, st));
    // large dynamic parts: the decoder is queued inside the pipeline (after the scan) and scatters its own terms' digits
    // (not for sources that go through the staging ring: there the calling thread copies while the decoder runs, and a
    // decoder that has to wait for every scalar first exposes that copy: 3.64 against 3.45 ms/step from pageable memory)
    const bool fuse = !use_pc && fuse_scatter_pays(n_dyn) && !would_stage(ctx, points_dyn32_host, n_dyn * 32) &&
                      !would_stage(ctx, scalars_dyn32_host, n_dyn * 32);
    FusedDyn fd; fd.points32_host = points_dyn32_host; fd.n = n_dyn;
    if (!fuse) TRY(start_upload_decode(ctx, points_dyn32_host, n_dyn, nullptr, 0, nullptr));
    size_t pos = 0;
    for (int k = 0; k < npieces; k++) {
        TRY(h2d(ctx, (uint8_t*)ctx->scalars.p + pos, pieces[k].host, pieces[k].bytes, st));
        pos += pieces[k].bytes;
    }
    TRY(h2d(ctx, (uint8_t*)ctx->scalars.p + n_static * 32, scalars_dyn32_host, n_dyn * 32, st));
    // fused route: the decoder's first chunk needs EVERY scalar (the histogram), so the scalars cross the link first and
    // alone; the side streams' point copies fork off behind them instead of sharing the link with them
    if (fuse) CK(ctx, cudaEventRecord(ctx->ev_fork, st));
    const uint4* ta = n_static ? t->d + offset * 6 : (const uint4*)ctx->dyn_table.p;
    TRY(msm_enqueue(ctx, plan, ctx->scalars.p, ta, (const uint4*)ctx->dyn_table.p, n_static, ctx->out_ext.p, nullptr, false,
                    fuse_encode ? ctx->out32.p : nullptr, fuse ? &fd : nullptr));
    if (ctx->profiling && n == 0) for (int i = 1; i < 4; i++) CK(ctx, cudaEventRecord(ctx->ev[i], st));
    CK(ctx, cudaMemcpyAsync(ctx->h_out + 32, ctx->bad.p, 8, cudaMemcpyDeviceToHost, st));    // after the join inside the pipeline
    return ZK_OK;
}

static int enqueue_partial_checked(zk_ctx* ctx, const zk_host_piece* pieces, int npieces, const zk_table* t, size_t offset, size_t n_static,
                                   const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host, size_t n_dyn, bool fuse_encode) {
    if (!ctx) return ZK_ERR_ARG;
    size_t sb = 0;
    for (int k = 0; k < npieces; k++) { if (pieces[k].bytes && !pieces[k].host) return ZK_ERR_ARG; sb += pieces[k].bytes; }
    if (sb != n_static * 32) return ZK_ERR_ARG;
    if (n_static && (!t || offset > t->len || n_static > t->len - offset)) return ZK_ERR_ARG;
    if (n_dyn && (!scalars_dyn32_host || !points_dyn32_host)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    int rc = enqueue_partial_impl(ctx, pieces, npieces, t, offset, n_static, scalars_dyn32_host, points_dyn32_host, n_dyn, fuse_encode);
    if (rc != ZK_OK) quiesce(ctx);
    return rc;
}
int zk_internal_enqueue_partial(zk_ctx* ctx, const zk_host_piece* pieces, int npieces, const zk_table* t, size_t offset, size_t n_static,
                                const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host, size_t n_dyn) {
    return enqueue_partial_checked(ctx, pieces, npieces, t, offset, n_static, scalars_dyn32_host, points_dyn32_host, n_dyn, false);
}
void* zk_internal_partial_ptr(zk_ctx* ctx) { return ctx->out_ext.p; }
int zk_internal_finish_partial(zk_ctx* ctx, size_t* bad_index) {
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, wait_main(ctx));
    unsigned long long b; memcpy(&b, ctx->h_out + 32, 8);
    if (bad_index) *bad_index = b == ~0ull ? (size_t)-1 : (size_t)b;
    return b == ~0ull ? ZK_OK : ZK_ERR_INVALID_POINT;
}

extern "C" int zk_msm_vartime_mixed(zk_ctx* ctx, const uint8_t* scalars_static32_host, const zk_table* t, size_t offset, size_t n_static,
                                    const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host, size_t n_dyn, uint8_t out32[32]) {
    if (!ctx || !out32) return ZK_ERR_ARG;
    if (n_static && !scalars_static32_host) return ZK_ERR_ARG;
    zk_host_piece piece = {scalars_static32_host, n_static * 32};
    TRY(enqueue_partial_checked(ctx, &piece, 1, t, offset, n_static, scalars_dyn32_host, points_dyn32_host, n_dyn, true));
    int rc = finish_encode(ctx, nullptr, 1, out32);
    if (rc != ZK_OK) { quiesce(ctx); return rc; }
    unsigned long long b; memcpy(&b, ctx->h_out + 32, 8);
    if (n_dyn && b != ~0ull) { memset(out32, 0, 32); return ZK_ERR_INVALID_POINT; }
    return ZK_OK;
}

extern "C" int zk_msm_vartime_table(zk_ctx* ctx, const uint8_t* scalars32_host, const zk_table* t, size_t offset, size_t n, uint8_t out32[32]) {
    if (!t) return ZK_ERR_ARG;
    return zk_msm_vartime_mixed(ctx, scalars32_host, t, offset, n, nullptr, nullptr, 0, out32);
}

extern "C" int zk_msm_vartime(zk_ctx* ctx, const uint8_t* scalars32_host, const uint8_t* points32_host, size_t n, uint8_t out32[32]) {
    return zk_msm_vartime_mixed(ctx, nullptr, nullptr, 0, 0, scalars32_host, points32_host, n, out32);
}

// ---- batches of independent MSMs -------------------------------------------------------------------
static int batch_impl(zk_ctx* ctx, const uint8_t* scalars32_host, const uint8_t* points32_host, const zk_table* t, size_t offset,
                      const uint64_t* seg_offsets, size_t m, size_t n, size_t longest, uint8_t* out32s, uint8_t* valid,
                      uint32_t* h_seg, uint32_t* h_bad) {
    cudaStream_t st = ctx->stream;
    Precomp pc;
    const bool use_pc = t && table_precomp(t, offset, longest, &pc);
    MsmPlan plan;
    TRY(msm_plan(ctx, n, m, use_pc ? &pc : nullptr, &plan));
    TRY(ensure(ctx, ctx->scalars, n * 32));
    TRY(ensure(ctx, ctx->seg, (m + 1) * 4 + m * 4));          // offsets, then per-MSM reject flags
    TRY(ensure(ctx, ctx->batch_ext, m * 128));
    TRY(ensure(ctx, ctx->batch_out, m * 32));
    TRY(ensure(ctx, ctx->bad, 8));
    if (!t) {
        TRY(ensure(ctx, ctx->comp, n * 32));
        TRY(ensure(ctx, ctx->dyn_table, n * 96));
    }
    for (size_t i = 0; i <= m; i++) h_seg[i] = (uint32_t)seg_offsets[i];
    uint32_t* seg_dev = (uint32_t*)ctx->seg.p;
    uint32_t* bad_msm_dev = seg_dev + (m + 1);
    CK(ctx, cudaMemcpyAsync(seg_dev, h_seg, (m + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaStreamSynchronize(st));                       // h_seg is pageable: the copy is done before anything forks
    CK(ctx, cudaMemsetAsync(bad_msm_dev, 0, m * 4, st));
    CK(ctx, cudaMemsetAsync(ctx->bad.p, 0xff, 8, st));
    const uint4* tab;
    if (t) tab = t->d + offset * 6;
    else {
        TRY(start_upload_decode(ctx, points32_host, n, seg_dev, (uint32_t)m, bad_msm_dev));
        tab = (const uint4*)ctx->dyn_table.p;
    }
    TRY(h2d(ctx, ctx->scalars.p, scalars32_host, n * 32, st));
    TRY(msm_enqueue(ctx, plan, ctx->scalars.p, tab, tab, (size_t)0xffffffffu, ctx->batch_ext.p, seg_dev, t != nullptr));
    k_encode_batch<<<grid_for(m, 64), 64, 0, st>>>((const uint4*)ctx->batch_ext.p, m, (uint4*)ctx->batch_out.p);
    LAUNCH_CHECK(ctx);
    if (ctx->profiling) CK(ctx, cudaEventRecord(ctx->ev[4], st));
    CK(ctx, cudaMemcpyAsync(out32s, ctx->batch_out.p, m * 32, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaMemcpyAsync(h_bad, bad_msm_dev, m * 4, cudaMemcpyDeviceToHost, st));
    CK(ctx, wait_main(ctx));
    if (ctx->profiling) {
        for (int i = 1; i < 4; i++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, ctx->ev[i], ctx->ev[i + 1]) == cudaSuccess) ctx->phase_ms[i] = ms;
        }
        ctx->phase_ms[0] = 0;
    }
    int rc = ZK_OK;
    for (size_t i = 0; i < m; i++) {
        if (h_bad[i]) { memset(out32s + 32 * i, 0, 32); rc = ZK_ERR_INVALID_POINT; }
        if (valid) valid[i] = h_bad[i] ? 0 : 1;
    }
    return rc;
}

static int batch_common(zk_ctx* ctx, const uint8_t* scalars32_host, const uint8_t* points32_host, const zk_table* t, size_t offset,
                        const uint64_t* seg_offsets, size_t m, uint8_t* out32s, uint8_t* valid) {
    if (!ctx || !seg_offsets || !out32s || m == 0 || seg_offsets[0] != 0) return ZK_ERR_ARG;
    const size_t n = (size_t)seg_offsets[m];
    if (n >= (1ull << 31) || m >= (1ull << 24)) return ZK_ERR_ARG;
    size_t longest = 0;
    for (size_t i = 0; i < m; i++) {
        if (seg_offsets[i + 1] < seg_offsets[i]) return ZK_ERR_ARG;
        size_t len = (size_t)(seg_offsets[i + 1] - seg_offsets[i]);
        if (len > longest) longest = len;
    }
    if (n && !scalars32_host) return ZK_ERR_ARG;
    if (t ? (offset > t->len || longest > t->len - offset) : (n && !points32_host)) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    uint32_t* h_seg = (uint32_t*)malloc((m + 1) * 4);
    uint32_t* h_bad = (uint32_t*)calloc(m, 4);
    int rc = (h_seg && h_bad) ? batch_impl(ctx, scalars32_host, points32_host, t, offset, seg_offsets, m, n, longest, out32s, valid, h_seg, h_bad)
                              : ZK_ERR_NOMEM;
    if (rc != ZK_OK && rc != ZK_ERR_INVALID_POINT) quiesce(ctx);    // nothing may still be reading h_seg / writing h_bad
    free(h_seg); free(h_bad);
    return rc;
}

extern "C" int zk_msm_vartime_batch(zk_ctx* ctx, const uint8_t* scalars32_host, const uint8_t* points32_host,
                                    const uint64_t* seg_offsets, size_t m, uint8_t* out32s, uint8_t* valid) {
    return batch_common(ctx, scalars32_host, points32_host, nullptr, 0, seg_offsets, m, out32s, valid);
}
extern "C" int zk_msm_vartime_table_batch(zk_ctx* ctx, const uint8_t* scalars32_host, const zk_table* t, size_t offset,
                                          const uint64_t* seg_offsets, size_t m, uint8_t* out32s) {
    if (!t) return ZK_ERR_ARG;
    return batch_common(ctx, scalars32_host, nullptr, t, offset, seg_offsets, m, out32s, nullptr);
}

extern "C" int zk_sum_compressed(zk_ctx* ctx, const uint8_t* points32_host, size_t g, uint8_t out32[32]) {
    if (!ctx || !out32 || (g && !points32_host) || g > 1024) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->comp, g * 32 + 32));
    TRY(ensure(ctx, ctx->out32, 32));
    TRY(ensure(ctx, ctx->bad, 8));
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemsetAsync(ctx->bad.p, 0xff, 8, st));
    if (g) CK(ctx, cudaMemcpyAsync(ctx->comp.p, points32_host, g * 32, cudaMemcpyHostToDevice, st));
    k_sum_compressed<<<1, 256, 0, st>>>((const uint4*)ctx->comp.p, (uint32_t)g, (uint4*)ctx->out32.p, (unsigned long long*)ctx->bad.p);
    LAUNCH_CHECK(ctx);
    CK(ctx, cudaMemcpyAsync(ctx->h_out, ctx->out32.p, 32, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaMemcpyAsync(ctx->h_out + 32, ctx->bad.p, 8, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    unsigned long long b; memcpy(&b, ctx->h_out + 32, 8);
    if (b != ~0ull) { memset(out32, 0, 32); return ZK_ERR_INVALID_POINT; }
    memcpy(out32, ctx->h_out, 32);
    return ZK_OK;
}

// Recreate the ctx's streams with the device's highest (1) or default (0) priority.  Call before queueing work.
extern "C" int zk_ctx_set_priority(zk_ctx* ctx, int high) {
    if (!ctx) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    int lo = 0, hi = 0;
    CK(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int prio = high ? hi : 0;
    cudaStream_t* all[3] = {&ctx->stream, &ctx->aux[0], &ctx->aux[1]};
    for (cudaStream_t* sp : all) {
        CK(ctx, cudaStreamSynchronize(*sp));
        cudaStream_t ns = nullptr;
        CK(ctx, cudaStreamCreateWithPriority(&ns, cudaStreamNonBlocking, prio));
        CK(ctx, cudaStreamDestroy(*sp));
        *sp = ns;
    }
    return ZK_OK;
}

extern "C" int zk_encoding_is_identity(const uint8_t enc32[32]) {
    if (!enc32) return 0;
    uint8_t o = 0;
    for (int i = 0; i < 32; i++) o |= enc32[i];
    return o == 0;
}

extern "C" int zk_bench_int_pipe(zk_ctx* ctx, int kind, double* ops_per_sec) {
    if (!ctx || !ops_per_sec || kind < 0 || kind > 3) return ZK_ERR_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    TRY(ensure(ctx, ctx->out32, 32));
    const int blocks = (ctx->sm_count > 0 ? ctx->sm_count : 148) * 8, threads = 256;
    const int iters = kind <= 1 ? 4096 : 512;
    double per_thread = kind == 0 ? 8.0 * iters : kind == 1 ? 8.0 * iters : 2.0 * iters;
    cudaEvent_t a = ctx->ev[0], b = ctx->ev[1];
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(ctx, cudaEventRecord(a, ctx->stream));
        if (kind == 0) k_bench_imad_wide<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->out32.p, iters, 12345u + rep);
        else if (kind == 1) k_bench_imad32<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->out32.p, iters, 12345u + rep);
        else k_bench_fe<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->out32.p, iters, 12345u + rep, kind == 3);
        LAUNCH_CHECK(ctx);
        CK(ctx, cudaEventRecord(b, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0; CK(ctx, cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    *ops_per_sec = per_thread * blocks * threads / (best * 1e-3);
    return ZK_OK;
}
