/* zkmsm.h -- C ABI of the B200 (sm_100a) Ristretto255 variable-time multiscalar-multiplication backend.
 *
 * This is the drop-in boundary for ZkVM's one data-parallel hot path: the large vartime MSM
 * that bulletproofs' r1cs::Verifier and the Schnorr/MuSig BatchVerifier collapse into, i.e. what
 * curve25519-dalek exposes as `VartimeMultiscalarMul::{vartime_multiscalar_mul,
 * optional_multiscalar_mul}` for `RistrettoPoint`.
 *
 * REFERENCE CITATIONS: none are possible.  /root/reference contains only README.md:1-7 (a
 * "repository has moved" notice) and license.txt:1-201.  The dalek trait
 * names above are recalled from public knowledge of the upstream crates (SURVEY.md Appendix A.1,
 * "UNVERIFIED RECALL"), not read from a mounted file.  Each entry point below therefore cites
 * the public standard that fixes its behaviour (RFC 9496) and names the dalek item it is meant to
 * stand behind; INTEGRATION.md shows the Rust `extern "C"` stub a maintainer would add.
 *
 * Conventions: every function returns ZK_OK (0) or a negative zk_status.  No exceptions, no
 * unwinding, no global state besides the CUDA primary context.  All buffers are caller-owned
 * unless documented.  `*_host` pointers are ordinary host memory (pinned memory makes the
 * copies faster but is not required); `*_dev` pointers are device pointers valid on the
 * context's device (e.g. torch tensors' data_ptr()).  A zk_ctx is not thread-safe: use one per
 * thread (they share the device).  Scalars are 32-byte little-endian integers; any 256-bit
 * value is accepted and is used modulo the group order l (identical result bytes, because the
 * ristretto255 group has prime order l).  Points are 32-byte ristretto255 encodings (RFC 9496
 * section 4.3.1).  There is no CPU fallback: without a CUDA device every call fails with
 * ZK_ERR_CUDA.
 */
#ifndef ZKMSM_H
#define ZKMSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum zk_status {
    ZK_OK = 0,
    ZK_ERR_CUDA = -1,          /* a CUDA runtime call failed; see zk_last_error() */
    ZK_ERR_INVALID_POINT = -2, /* some encoding failed RFC 9496 4.3.1 Decode: the `None` of optional_multiscalar_mul */
    ZK_ERR_ARG = -3,           /* null pointer / size mismatch / out-of-range slice */
    ZK_ERR_NOMEM = -4
} zk_status;

typedef struct zk_ctx zk_ctx;     /* per-thread handle: device, stream, reusable workspace */
typedef struct zk_table zk_table; /* device-resident decompressed point cache (affine Niels, 96 B/point) */
typedef struct zk_mgpu zk_mgpu;   /* one handle over several GPUs of a box: one zk_ctx + one worker thread per device */
typedef struct zk_mgpu_table zk_mgpu_table; /* point cache sharded by index range over the devices of a zk_mgpu */

#define ZK_ABI_VERSION 2
int zk_abi_version(void);
const char* zk_status_str(int status);
/* Text of the last CUDA failure seen by this ctx (empty string if none). */
const char* zk_last_error(const zk_ctx* ctx);

/* ---- context ---- */
int zk_ctx_create(int device, zk_ctx** out);
void zk_ctx_destroy(zk_ctx* ctx);
/* Blocks until all work queued by this ctx has finished. */
int zk_ctx_sync(zk_ctx* ctx);
/* How a call waits for the device: 0 = spin (default, lowest latency), 1 = park the calling thread on a blocking-sync
 * event (about 0.3 ms more latency per call, no CPU while waiting).  Use 1 when more calls are in flight on the host
 * than it has cores to spare (several contexts per GPU x several GPUs): a spinning waiter occupies a core. */
int zk_ctx_set_wait(zk_ctx* ctx, int mode);
/* cudaStream_t of the ctx as an opaque pointer: the MAIN stream, on which every asynchronous entry point (zk_*_dev) is
 * ordered -- work queued on it afterwards sees the call's result, events recorded on it bracket the call.  Internally a
 * call also uses two decode side streams and a high-priority stream for its short latency-bound phases; the main stream
 * waits for them, so callers never see them.  The main stream has low priority: latency-critical follow-up work (e.g. a
 * 128-byte gather of partials) is better issued on a high-priority stream of the caller's that waits on this one. */
void* zk_ctx_stream(zk_ctx* ctx);
/* high = 1: recreate the ctx's streams with the device's highest stream priority, so that its (small) kernels are
 * scheduled ahead of other contexts' bulk work; 0 restores the default.  Call while the ctx is idle. */
int zk_ctx_set_priority(zk_ctx* ctx, int high);

/* ---- host buffers ----
 * Every `*_host` pointer may be ordinary PAGEABLE memory (a Rust Vec<u8>).  Uploads of 256 KiB or more from pageable
 * memory go through a pinned staging ring inside the ctx (4 x 4 MiB): the calling thread copies chunk i+1 while the DMA
 * engine moves chunk i and the device decodes chunk i-1, and the caller's buffer may be reused as soon as the call
 * returns.  Page-locked sources (cudaHostAlloc / zk_host_register) are detected and copied directly, which is faster:
 * register long-lived buffers once with zk_host_register().  zk_ctx_set_staging: 0 = auto (default), 1 = never stage
 * (plain cudaMemcpyAsync, which blocks on pageable sources), 2 = always stage. */
int zk_host_register(void* ptr, size_t bytes);
int zk_host_unregister(void* ptr);
int zk_ctx_set_staging(zk_ctx* ctx, int mode);
/* Bytes this ctx has moved through its staging ring so far. */
uint64_t zk_ctx_staged_bytes(const zk_ctx* ctx);

/* ---- point tables: "decompress once, cache on device" ----
 * Stand behind: CompressedRistretto::decompress (RFC 9496 4.3.1), RistrettoPoint::from_uniform_bytes
 * (RFC 9496 4.3.4; what bulletproofs' generator chains call), and holding Vec<RistrettoPoint> of
 * static generators (BulletproofGens / PedersenGens) across verifications. */
int zk_table_create(zk_ctx* ctx, size_t capacity, zk_table** out);
void zk_table_destroy(zk_table* t);
size_t zk_table_len(const zk_table* t);
size_t zk_table_capacity(const zk_table* t);
/* Sets len = 0 (keeps the allocation). */
void zk_table_clear(zk_table* t);
/* Decode n encodings and append them.  On ZK_ERR_INVALID_POINT nothing is appended and, if
 * bad_index != NULL, *bad_index = the lowest failing index. */
int zk_table_append_compressed(zk_ctx* ctx, zk_table* t, const uint8_t* points32_host, size_t n, size_t* bad_index);
int zk_table_append_compressed_dev(zk_ctx* ctx, zk_table* t, const void* points32_dev, size_t n, size_t* bad_index);
/* Hash-to-group (RFC 9496 4.3.4) of n 64-byte strings, appended. */
int zk_table_append_uniform(zk_ctx* ctx, zk_table* t, const uint8_t* bytes64_host, size_t n);
int zk_table_append_uniform_dev(zk_ctx* ctx, zk_table* t, const void* bytes64_dev, size_t n);
/* Append n already-decompressed points given in extended coordinates: X, Y, Z, T as 4 x 32-byte canonical
 * little-endian field elements (128 bytes per point; any projective representative with Z != 0).  Stands behind
 * "the caller holds Vec<RistrettoPoint>" (dalek's RistrettoPoint is four field elements; FieldElement::to_bytes
 * gives this form without the inversion + square root a CPU-side compress() would cost).  Normalised to Z = 1 on the
 * device.  Fully validated: ZK_ERR_INVALID_POINT (+ lowest index) for a non-canonical coordinate, Z = 0, an off-curve
 * point, T*Z != X*Y, or a point outside the even subgroup 2E (ristretto255 = 2E / E[4], RFC 9496 section 3: a point with
 * an odd 8-torsion component is not the representative of any element); nothing is appended then. */
int zk_table_append_extended(zk_ctx* ctx, zk_table* t, const uint8_t* ext128_host, size_t n, size_t* bad_index);
int zk_table_append_extended_dev(zk_ctx* ctx, zk_table* t, const void* ext128_dev, size_t n, size_t* bad_index);
/* Same input, for points the caller KNOWS to be valid representatives (values of dalek's RistrettoPoint are by
 * construction): coordinates are taken modulo p, only Z = 0 is rejected, and the normalisation uses one batched
 * inversion per 8 points instead of one exponentiation per point (about 7x faster).  Passing anything that is not a
 * valid representative gives undefined results -- use the checked form for untrusted input. */
int zk_table_append_extended_unchecked(zk_ctx* ctx, zk_table* t, const uint8_t* ext128_host, size_t n, size_t* bad_index);
int zk_table_append_extended_unchecked_dev(zk_ctx* ctx, zk_table* t, const void* ext128_dev, size_t n, size_t* bad_index);
/* Optional, for STATIC generator sets that are multiplied against many times: expand the table into per-window
 * multiples 2^(c*w) * P (w = 0 .. ceil(253/c)), so that an MSM over it needs no doublings at all and every window shares
 * one bucket set (fewer additions per point: the width can grow without growing the tree).  Costs W x the memory
 * (W * 96 B per point) and ~253 doublings + W inversions per point, once.  c = 0 picks the width from the table
 * length; 4 <= c <= 20 otherwise.  zk_msm_vartime_table / _table_batch / zk_msm_table_dev use the expansion
 * automatically when it pays for the slice at hand (a short slice of a long table takes the plain route; while the
 * expansion is in use its width overrides zk_ctx_set_window); a successful append to the table, or clearing it, drops
 * it.  Results are bit-identical either way.
 * (The role dalek's VartimePrecomputedMultiscalarMul plays for static points.) */
int zk_table_precompute(zk_ctx* ctx, zk_table* t, int c);
/* Window width of the table's expansion, 0 if it has none. */
int zk_table_precomputed_window(const zk_table* t);
/* Encode table[offset .. offset+n) (RFC 9496 4.3.2) into out32 (n*32 bytes). */
int zk_table_compress(zk_ctx* ctx, const zk_table* t, size_t offset, size_t n, uint8_t* out32_host);
int zk_table_compress_dev(zk_ctx* ctx, const zk_table* t, size_t offset, size_t n, void* out32_dev);

/* ---- the hot path ----
 * out32 = Encode( sum_i scalars[i] * Decode(points[i]) ).
 * Stands behind RistrettoPoint::optional_multiscalar_mul(scalars, points.map(decompress)) followed
 * by .compress(): ZK_ERR_INVALID_POINT <=> None.  n == 0 yields the identity encoding (32 zero bytes). */
int zk_msm_vartime(zk_ctx* ctx, const uint8_t* scalars32_host, const uint8_t* points32_host, size_t n,
                   uint8_t out32[32]);

/* Same, over cached points table[offset .. offset+n).  Stands behind
 * RistrettoPoint::vartime_multiscalar_mul(scalars, &gens[offset..offset+n]).compress(). */
int zk_msm_vartime_table(zk_ctx* ctx, const uint8_t* scalars32_host, const zk_table* t, size_t offset, size_t n,
                         uint8_t out32[32]);

/* Static-prefix + dynamic-suffix form (the shape of a bulletproofs verification MSM: cached
 * generators first, then the proof's own compressed points):
 *   sum_{i<n_static} s_static[i]*table[offset+i]  +  sum_{j<n_dyn} s_dyn[j]*Decode(points_dyn[j]). */
int zk_msm_vartime_mixed(zk_ctx* ctx, const uint8_t* scalars_static32_host, const zk_table* t, size_t offset,
                         size_t n_static, const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host,
                         size_t n_dyn, uint8_t out32[32]);

/* ---- batches of independent MSMs (one verdict per proof) ----
 * m MSMs in ONE pass over the device: MSM k covers terms [seg_offsets[k], seg_offsets[k+1]) of the concatenated
 * scalar/point arrays (seg_offsets has m+1 entries, seg_offsets[0] = 0).  out32s receives m encodings.  Small MSMs
 * are latency-bound one at a time (253 dependent doublings in the window Horner); batched, they share every kernel
 * launch and hide each other's serial tails.  This is the shape of "verify 1024 transactions, each with its own
 * accept/reject": unlike folding all proofs into one MSM with random weights, one bad proof does not void the rest.
 * An invalid encoding voids only its own MSM: valid[k] = 0 (if valid != NULL), out32s[k] = 0, and the call returns
 * ZK_ERR_INVALID_POINT; all other results are still written. */
int zk_msm_vartime_batch(zk_ctx* ctx, const uint8_t* scalars32_host, const uint8_t* points32_host,
                         const uint64_t* seg_offsets, size_t m, uint8_t* out32s, uint8_t* valid);
/* Same, every MSM running over the SAME cached points: MSM k = sum_j scalars[seg_offsets[k] + j] * table[offset + j]
 * (m proofs verified against one set of static generators). */
int zk_msm_vartime_table_batch(zk_ctx* ctx, const uint8_t* scalars32_host, const zk_table* t, size_t offset,
                               const uint64_t* seg_offsets, size_t m, uint8_t* out32s);

/* Device-resident form: scalars already in HBM (n*32 bytes), result left in HBM as an extended
 * point (X,Y,Z,T: 4 x 32-byte little-endian field elements, 128 bytes) so that per-GPU partial
 * sums can be gathered with one collective.  Asynchronous on the ctx stream. */
int zk_msm_table_dev(zk_ctx* ctx, const void* scalars32_dev, const zk_table* t, size_t offset, size_t n,
                     void* out_ext128_dev);
/* Sum g extended points (g*128 bytes in HBM, e.g. the all-gathered partials) and encode. */
int zk_ext_sum_compress_dev(zk_ctx* ctx, const void* ext128_dev, size_t g, uint8_t out32[32]);
/* out32 = Encode(sum_i Decode(points[i])), g <= 1024, in ONE small kernel launch: the combine step of a deployment
 * with one process per GPU, where every process returns the 32-byte encoding of its partial MSM
 * (the sum of RistrettoPoints, `iter.sum()`).  ZK_ERR_INVALID_POINT if an encoding is rejected. */
int zk_sum_compressed(zk_ctx* ctx, const uint8_t* points32_host, size_t g, uint8_t out32[32]);
/* Is the ristretto255 element the identity?  (the accept test of both verifiers) */
int zk_encoding_is_identity(const uint8_t enc32[32]);

/* ---- several GPUs behind one call (BASELINE.json config 5: point-range shards + one gather of partial points) ----
 * devices = g distinct CUDA device ordinals (NULL: 0 .. g-1).  Device r runs the whole pipeline over the index range
 * [r*n/g, (r+1)*n/g) of the call's terms (remainder spread over the first ranks) and leaves a 128-byte extended partial
 * in its HBM; the g partials are gathered on devices[0], added and encoded there.  Gather: 0 = g-1 peer copies of
 * 128 bytes (cudaMemcpyPeerAsync over NVLink; default), 1 = one ncclAllGather on a communicator from ncclCommInitAll
 * (libnccl.so.2 is loaded at run time; ZK_ERR_CUDA if it cannot be).  A zk_mgpu is not thread-safe: one per thread.
 * Stand behind the same dalek calls as the single-GPU entry points; results are bit-identical to them. */
int zk_mgpu_create(const int* devices, int g, zk_mgpu** out);
void zk_mgpu_destroy(zk_mgpu* mg);
int zk_mgpu_device_count(const zk_mgpu* mg);
const char* zk_mgpu_last_error(const zk_mgpu* mg);
int zk_mgpu_set_gather(zk_mgpu* mg, int mode);
int zk_mgpu_set_staging(zk_mgpu* mg, int mode);
int zk_mgpu_set_wait(zk_mgpu* mg, int mode);
uint64_t zk_mgpu_launch_count(const zk_mgpu* mg);
/* out32 = Encode(sum_i scalars[i] * Decode(points[i])) over all g devices; ZK_ERR_INVALID_POINT <=> None. */
int zk_mgpu_msm_vartime(zk_mgpu* mg, const uint8_t* scalars32_host, const uint8_t* points32_host, size_t n,
                        uint8_t out32[32]);
/* Sharded point cache: every append is cut into g index ranges, device r keeps its range. */
int zk_mgpu_table_create(zk_mgpu* mg, size_t capacity, zk_mgpu_table** out);
void zk_mgpu_table_destroy(zk_mgpu_table* t);
size_t zk_mgpu_table_len(const zk_mgpu_table* t);
void zk_mgpu_table_clear(zk_mgpu_table* t);
/* All-or-nothing; *bad_index = lowest failing index of THIS append. */
int zk_mgpu_table_append_compressed(zk_mgpu_table* t, const uint8_t* points32_host, size_t n, size_t* bad_index);
int zk_mgpu_table_append_uniform(zk_mgpu_table* t, const uint8_t* bytes64_host, size_t n);
/* sum_i scalars[i] * table[offset + i], i < n, over all g devices. */
int zk_mgpu_msm_vartime_table(zk_mgpu* mg, const uint8_t* scalars32_host, const zk_mgpu_table* t, size_t offset,
                              size_t n, uint8_t out32[32]);

/* Static prefix from the sharded cache + dynamic compressed suffix (zk_msm_vartime_mixed over all g devices). */
int zk_mgpu_msm_vartime_mixed(zk_mgpu* mg, const uint8_t* scalars_static32_host, const zk_mgpu_table* t, size_t offset,
                              size_t n_static, const uint8_t* scalars_dyn32_host, const uint8_t* points_dyn32_host,
                              size_t n_dyn, uint8_t out32[32]);

/* ---- tuning / measurement ---- */
/* Force the Pippenger window width (bits, 4..20); 0 restores the size-based choice. */
int zk_ctx_set_window(zk_ctx* ctx, int c);
/* Window the size-based rule picks for an n-point MSM. */
int zk_pick_window(size_t n);
/* Integer-pipe microbenchmark: sustained 32x32+64 multiply-accumulate (IMAD.WIDE.U32 with carry)
 * lane-operations per second on this device -- the denominator of the IMAD roofline.
 * kind: 0 = IMAD.WIDE.U32 carry chains, 1 = plain 32-bit IMAD, 2 = field multiplies/s (fe_mul), 3 = field squarings/s. */
int zk_bench_int_pipe(zk_ctx* ctx, int kind, double* ops_per_sec);
/* Per-phase device time of the most recent MSM on this ctx, in milliseconds (CUDA events):
 * [0] decompress, [1] digit histogram + scan + scatter, [2] bucket accumulation,
 * [3] bucket/window reduction + encode.  Only recorded after zk_ctx_set_profiling(ctx, 1). */
int zk_ctx_set_profiling(zk_ctx* ctx, int on);
int zk_ctx_last_phase_ms(zk_ctx* ctx, float out_ms[4]);
/* Number of kernel launches issued by this ctx so far. */
uint64_t zk_ctx_launch_count(const zk_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ZKMSM_H */
