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

#define ZK_ABI_VERSION 1
int zk_abi_version(void);
const char* zk_status_str(int status);
/* Text of the last CUDA failure seen by this ctx (empty string if none). */
const char* zk_last_error(const zk_ctx* ctx);

/* ---- context ---- */
int zk_ctx_create(int device, zk_ctx** out);
void zk_ctx_destroy(zk_ctx* ctx);
/* Blocks until all work queued by this ctx has finished. */
int zk_ctx_sync(zk_ctx* ctx);
/* cudaStream_t of the ctx as an opaque pointer (for CUDA-event timing by the caller). */
void* zk_ctx_stream(zk_ctx* ctx);

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
 * device.  ZK_ERR_INVALID_POINT (+ lowest index) for a non-canonical coordinate, Z = 0, an off-curve point or
 * T*Z != X*Y; nothing is appended then. */
int zk_table_append_extended(zk_ctx* ctx, zk_table* t, const uint8_t* ext128_host, size_t n, size_t* bad_index);
int zk_table_append_extended_dev(zk_ctx* ctx, zk_table* t, const void* ext128_dev, size_t n, size_t* bad_index);
/* Optional, for STATIC generator sets that are multiplied against many times: expand the table into per-window
 * multiples 2^(c*w) * P (w = 0 .. ceil(253/c)), so that an MSM over it needs no doublings at all and every window shares
 * one bucket set (fewer additions per point: the width can grow without growing the tree).  Costs W x the memory
 * (W * 96 B per point) and ~253 doublings + W inversions per point, once.  c = 0 picks the width from the table
 * length; 4 <= c <= 20 otherwise.  zk_msm_vartime_table / _table_batch / zk_msm_table_dev use the expansion
 * automatically; appending to the table or clearing it drops it.  Results are bit-identical either way.
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
/* Is the ristretto255 element the identity?  (the accept test of both verifiers) */
int zk_encoding_is_identity(const uint8_t enc32[32]);

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
