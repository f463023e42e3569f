// GF(2^255-19) arithmetic for sm_100a: 8 saturated 32-bit limbs, IMAD.WIDE carry chains.
//
// Representation: value v in [0, 2^256), little-endian limbs, only congruent to the field
// element mod p = 2^255-19 ("loose").  fe_freeze() gives the canonical representative.
// 2^256 = 38 (mod p) is the only reduction identity used on the hot path.
//
// Behavioural spec: RFC 9496 section 4.1/4.2 (field, constants, SQRT_RATIO_M1) and RFC 7748
// section 4.1 (the field).  The reference tree (/root/reference) contains no source for this
// path; nothing here is derived from it.
//
// The same header compiles for the host (tests/host_emul) with portable C in place of PTX, so
// the limb schedules can be unit-tested without a GPU.  The host rendering is test-only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__
#define ZK_INLINE __forceinline__
#else
#define ZK_HD
#define ZK_INLINE inline __attribute__((always_inline))
#endif

namespace zk {

struct fe { uint32_t v[8]; };

#include "fe25519_mul.inc"

// ---- constants (computed by tools/constants.py; cross-checked against RFC 9496 section 4.1) ----
#define ZK_FE_CONST(name, ...) \
    ZK_HD ZK_INLINE fe name() { fe r = {{__VA_ARGS__}}; return r; }
ZK_FE_CONST(fe_zero, 0, 0, 0, 0, 0, 0, 0, 0)
ZK_FE_CONST(fe_one, 1, 0, 0, 0, 0, 0, 0, 0)
ZK_FE_CONST(fe_d, 0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu)
ZK_FE_CONST(fe_d2, 0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu)
ZK_FE_CONST(fe_sqrt_m1, 0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u)
ZK_FE_CONST(fe_one_minus_d_sq, 0x945fc176u, 0xe27c09c1u, 0xcd5e350fu, 0x2c81a138u, 0xbe70dfe4u, 0x9994abddu, 0xb2b3e0d7u, 0x029072a8u)
ZK_FE_CONST(fe_d_minus_one_sq, 0x44ed4d20u, 0x31ad5aaau, 0xb01e1999u, 0xd29e4a2cu, 0x529b4eebu, 0x4cdcd32fu, 0xf66c2241u, 0x5968b37au)
ZK_FE_CONST(fe_invsqrt_a_minus_d, 0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u)
// RFC 9496 fixes the odd ("negative") root: 2506306895338462347411141415870215270124453150249265646007921048261043075023 5
ZK_FE_CONST(fe_sqrt_ad_minus_one, 0x497b2e1bu, 0x7e97f6a0u, 0x1b7854bdu, 0xaf9d8e0cu, 0x31f5d1fdu, 0x0f3cfcc9u, 0x2b8348acu, 0x376931bfu)

ZK_HD ZK_INLINE void fe_mul(fe& r, const fe& a, const fe& b) { fe_mul_limbs(r.v, a.v, b.v); }
ZK_HD ZK_INLINE void fe_sqr(fe& r, const fe& a) { fe_sqr_limbs(r.v, a.v); }

// r = a + b.  19 instructions: 8-limb carry chain, then the carry (weight 2^256 = 38) is folded twice.
ZK_HD ZK_INLINE void fe_add(fe& r, const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17; addc.cc.u32 %1, %10, %18; addc.cc.u32 %2, %11, %19; addc.cc.u32 %3, %12, %20;"
        "addc.cc.u32 %4, %13, %21; addc.cc.u32 %5, %14, %22; addc.cc.u32 %6, %15, %23; addc.cc.u32 %7, %16, %24;"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t m = c * 38u;
    asm("add.cc.u32 %0, %0, %9; addc.cc.u32 %1, %1, 0; addc.cc.u32 %2, %2, 0; addc.cc.u32 %3, %3, 0;"
        "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0;"
        "addc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=&r"(c)
        : "r"(m));
    r.v[0] += c * 38u;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    c *= 38u;
    for (int i = 0; i < 8; i++) { c += r.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    r.v[0] += (uint32_t)c * 38u;
#endif
}

// r = a - b.  A borrow means the result wrapped by +2^256 = +38 (mod p): take 38 back off, twice at most.
ZK_HD ZK_INLINE void fe_sub(fe& r, const fe& a, const fe& b) {
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17; subc.cc.u32 %1, %10, %18; subc.cc.u32 %2, %11, %19; subc.cc.u32 %3, %12, %20;"
        "subc.cc.u32 %4, %13, %21; subc.cc.u32 %5, %14, %22; subc.cc.u32 %6, %15, %23; subc.cc.u32 %7, %16, %24;"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t m = c & 38u;   // c is 0 or 0xffffffff
    asm("sub.cc.u32 %0, %0, %9; subc.cc.u32 %1, %1, 0; subc.cc.u32 %2, %2, 0; subc.cc.u32 %3, %3, 0;"
        "subc.cc.u32 %4, %4, 0; subc.cc.u32 %5, %5, 0; subc.cc.u32 %6, %6, 0; subc.cc.u32 %7, %7, 0;"
        "subc.u32 %8, 0, 0;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]),
          "+r"(r.v[7]), "=&r"(c)
        : "r"(m));
    r.v[0] -= c & 38u;
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - b.v[i]; r.v[i] = (uint32_t)c; c >>= 32; }
    int64_t m = c ? 38 : 0; c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)r.v[i] - (i == 0 ? m : 0); r.v[i] = (uint32_t)c; c >>= 32; }
    r.v[0] -= c ? 38u : 0u;
#endif
}

ZK_HD ZK_INLINE void fe_neg(fe& r, const fe& a) { fe z = fe_zero(); fe_sub(r, z, a); }
ZK_HD ZK_INLINE void fe_dbl(fe& r, const fe& a) { fe_add(r, a, a); }

// Canonical representative in [0, p).
ZK_HD ZK_INLINE void fe_freeze(fe& r, const fe& a) {
    // v < 2^256: fold bit 255 (2^255 = 19), giving v' < 2^255 + 19; then one conditional subtract of p.
    uint32_t t[8];
    uint64_t c = (uint64_t)(a.v[7] >> 31) * 19u;
    for (int i = 0; i < 8; i++) {
        c += (i == 7) ? (a.v[7] & 0x7fffffffu) : a.v[i];
        t[i] = (uint32_t)c; c >>= 32;
    }
    // u = t + 19; if u >= 2^255 then t >= p and the answer is u - 2^255, else t.
    uint32_t u[8];
    c = 19u;
    for (int i = 0; i < 8; i++) { c += t[i]; u[i] = (uint32_t)c; c >>= 32; }
    uint32_t ge = 0u - (u[7] >> 31);
    u[7] &= 0x7fffffffu;
    for (int i = 0; i < 8; i++) r.v[i] = (u[i] & ge) | (t[i] & ~ge);
}

ZK_HD ZK_INLINE bool fe_is_zero(const fe& a) {
    fe t; fe_freeze(t, a);
    uint32_t o = 0;
    for (int i = 0; i < 8; i++) o |= t.v[i];
    return o == 0;
}
ZK_HD ZK_INLINE bool fe_eq(const fe& a, const fe& b) { fe t; fe_sub(t, a, b); return fe_is_zero(t); }
// RFC 9496 IS_NEGATIVE: low bit of the canonical encoding.
ZK_HD ZK_INLINE bool fe_is_negative(const fe& a) { fe t; fe_freeze(t, a); return (t.v[0] & 1u) != 0; }
ZK_HD ZK_INLINE void fe_cneg(fe& r, const fe& a, bool neg) {
    fe n; fe_neg(n, a);
    for (int i = 0; i < 8; i++) r.v[i] = neg ? n.v[i] : a.v[i];
}
ZK_HD ZK_INLINE void fe_abs(fe& r, const fe& a) { fe_cneg(r, a, fe_is_negative(a)); }
ZK_HD ZK_INLINE void fe_select(fe& r, const fe& a, const fe& b, bool take_b) {
    for (int i = 0; i < 8; i++) r.v[i] = take_b ? b.v[i] : a.v[i];
}

// 32-byte little-endian decode.  Returns false when the encoding is not canonical (>= p or bit 255 set);
// r still holds the low 255 bits (RFC 9496 from_uniform_bytes masks bit 255 and reduces).
ZK_HD ZK_INLINE bool fe_from_words(fe& r, const uint32_t w[8]) {
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    bool high = (w[7] >> 31) != 0;
    r.v[7] &= 0x7fffffffu;
    // canonical iff value < p  <=>  value + 19 < 2^255
    uint64_t c = 19u;
    for (int i = 0; i < 8; i++) { c += r.v[i]; c = (i == 7) ? c : (c >> 32); }
    bool ge_p = (c >> 31) != 0;
    return !high && !ge_p;
}

#ifndef ZK_SQRN_UNROLL
#define ZK_SQRN_UNROLL 2   // measured on the decoder: 1.980 (1) / 1.939 (2) / 2.019 (4) ms per 2^20 points
#endif
constexpr int kSqrnUnroll = ZK_SQRN_UNROLL;
ZK_HD ZK_INLINE void fe_sqr_n(fe& r, const fe& a, int n) {
    fe_sqr(r, a);
#if defined(__CUDA_ARCH__)
#pragma unroll kSqrnUnroll
#endif
    for (int i = 1; i < n; i++) fe_sqr(r, r);
}

// Operation policies.  The throughput kernels inline every multiply (fe_ops_inline).  The single-warp tail kernels
// (window Horner, final Encode) execute a long stretch of straight-line field code exactly once: fully inlined it is
// hundreds of KB of SASS and the warp stalls on instruction fetch ("no_instruction" was the top stall of
// k_ext_sum_encode in round 1); fe_ops_call routes multiplies and squarings through three real functions instead, so
// the cold path is a few KB that stays in the instruction cache.
struct fe_ops_inline {
    static ZK_HD ZK_INLINE void mul(fe& r, const fe& a, const fe& b) { fe_mul(r, a, b); }
    static ZK_HD ZK_INLINE void sqr(fe& r, const fe& a) { fe_sqr(r, a); }
    static ZK_HD ZK_INLINE void sqr_n(fe& r, const fe& a, int n) { fe_sqr_n(r, a, n); }
};
#if defined(__CUDACC__)
// by value: 16 operand registers in, 8 out -- no stack traffic across the call
static __device__ __noinline__ fe fe_mul_call(fe a, fe b) { fe r; fe_mul(r, a, b); return r; }
static __device__ __noinline__ fe fe_sqr_n_call(fe a, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) fe_sqr(a, a);
    return a;
}
struct fe_ops_call {
    static __device__ __forceinline__ void mul(fe& r, const fe& a, const fe& b) { r = fe_mul_call(a, b); }
    static __device__ __forceinline__ void sqr(fe& r, const fe& a) { r = fe_sqr_n_call(a, 1); }
    static __device__ __forceinline__ void sqr_n(fe& r, const fe& a, int n) { r = fe_sqr_n_call(a, n); }
};
#endif

// Two field elements advanced in lockstep: the exponentiation chains below are 250 dependent squarings, so running
// two of them interleaved doubles the instruction-level parallelism a thread offers the IMAD pipe.
struct fe2 { fe a, b; };
ZK_HD ZK_INLINE void fe_mul(fe2& r, const fe2& x, const fe2& y) { fe_mul(r.a, x.a, y.a); fe_mul(r.b, x.b, y.b); }
ZK_HD ZK_INLINE void fe_sqr(fe2& r, const fe2& x) { fe_sqr(r.a, x.a); fe_sqr(r.b, x.b); }
ZK_HD ZK_INLINE void fe_sqr_n(fe2& r, const fe2& x, int n) {
    fe_sqr(r, x);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = 1; i < n; i++) fe_sqr(r, r);
}

// r = a^(2^252 - 3) = a^((p-5)/8).  Standard 2^k-1 ladder: 251 squarings + 11 multiplies.  F = fe or fe2.
template <class F>
ZK_HD inline void fe_pow22523_t(F& r, const F& z) {
    F t0, t1, t2;
    fe_sqr(t0, z);                 // 2
    fe_sqr_n(t1, t0, 2);           // 8
    fe_mul(t1, z, t1);             // 9
    fe_mul(t0, t0, t1);            // 11
    fe_sqr(t0, t0);                // 22
    fe_mul(t0, t1, t0);            // 31 = 2^5-1
    fe_sqr_n(t1, t0, 5);
    fe_mul(t0, t1, t0);            // 2^10-1
    fe_sqr_n(t1, t0, 10);
    fe_mul(t1, t1, t0);            // 2^20-1
    fe_sqr_n(t2, t1, 20);
    fe_mul(t1, t2, t1);            // 2^40-1
    fe_sqr_n(t1, t1, 10);
    fe_mul(t0, t1, t0);            // 2^50-1
    fe_sqr_n(t1, t0, 50);
    fe_mul(t1, t1, t0);            // 2^100-1
    fe_sqr_n(t2, t1, 100);
    fe_mul(t1, t2, t1);            // 2^200-1
    fe_sqr_n(t1, t1, 50);
    fe_mul(t0, t1, t0);            // 2^250-1
    fe_sqr_n(t0, t0, 2);           // 2^252-4
    fe_mul(r, t0, z);              // 2^252-3
}
ZK_HD inline void fe_pow22523(fe& r, const fe& z) { fe_pow22523_t<fe>(r, z); }
// Same ladder through an operation policy (see fe_ops_call).
template <class Ops>
ZK_HD inline void fe_pow22523_ops(fe& r, const fe& z) {
    fe t0, t1, t2;
    Ops::sqr(t0, z);
    Ops::sqr_n(t1, t0, 2);
    Ops::mul(t1, z, t1);
    Ops::mul(t0, t0, t1);
    Ops::sqr(t0, t0);
    Ops::mul(t0, t1, t0);            // 2^5-1
    Ops::sqr_n(t1, t0, 5);
    Ops::mul(t0, t1, t0);            // 2^10-1
    Ops::sqr_n(t1, t0, 10);
    Ops::mul(t1, t1, t0);            // 2^20-1
    Ops::sqr_n(t2, t1, 20);
    Ops::mul(t1, t2, t1);            // 2^40-1
    Ops::sqr_n(t1, t1, 10);
    Ops::mul(t0, t1, t0);            // 2^50-1
    Ops::sqr_n(t1, t0, 50);
    Ops::mul(t1, t1, t0);            // 2^100-1
    Ops::sqr_n(t2, t1, 100);
    Ops::mul(t1, t2, t1);            // 2^200-1
    Ops::sqr_n(t1, t1, 50);
    Ops::mul(t0, t1, t0);            // 2^250-1
    Ops::sqr_n(t0, t0, 2);           // 2^252-4
    Ops::mul(r, t0, z);              // 2^252-3
}

// r = a^(p-2) = a^(2^255-21).
ZK_HD inline void fe_invert(fe& r, const fe& z) {
    // a^(2^255-21) = (a^(2^252-3))^8 * a^3
    fe t, z3;
    fe_pow22523(t, z);
    fe_sqr_n(t, t, 3);             // a^(2^255-24)
    fe_sqr(z3, z); fe_mul(z3, z3, z);
    fe_mul(r, t, z3);
}

// RFC 9496 section 4.2 SQRT_RATIO_M1(u, v): returns was_square, r = |sqrt(u/v)| or |sqrt(i*u/v)|.
// Split around the exponentiation so that two instances can share one interleaved chain.
struct sqrt_ratio_state { fe v3, t; };
template <class Ops = fe_ops_inline>
ZK_HD ZK_INLINE void fe_sqrt_ratio_pre(sqrt_ratio_state& st, const fe& u, const fe& v) {
    fe v7;
    Ops::sqr(st.v3, v); Ops::mul(st.v3, st.v3, v);          // v^3
    Ops::sqr(v7, st.v3); Ops::mul(v7, v7, v);               // v^7
    Ops::mul(st.t, u, v7);                                  // the value to raise to (p-5)/8
}
template <class Ops = fe_ops_inline>
ZK_HD ZK_INLINE bool fe_sqrt_ratio_post(fe& r, const sqrt_ratio_state& st, const fe& powed, const fe& u, const fe& v) {
    fe t, check, nu, nui;
    Ops::mul(t, powed, st.v3); Ops::mul(t, t, u);           // r = u v^3 (u v^7)^((p-5)/8)
    Ops::sqr(check, t); Ops::mul(check, check, v);          // v r^2
    fe_neg(nu, u);
    fe i = fe_sqrt_m1();
    Ops::mul(nui, nu, i);
    bool correct = fe_eq(check, u);
    bool flipped = fe_eq(check, nu);
    bool flipped_i = fe_eq(check, nui);
    fe ri; Ops::mul(ri, t, i);
    fe_select(t, t, ri, flipped | flipped_i);
    fe_abs(r, t);
    return correct | flipped;
}
// SQRT_RATIO_M1(1, v), the form both Decode and Encode use: the general routine's three multiplications by u = 1 (and
// by the constant -i) are skipped.  Same outputs as fe_sqrt_ratio_m1(r, 1, v), bit for bit.
template <class Ops = fe_ops_inline>
ZK_HD inline bool fe_invsqrt(fe& r, const fe& v) {
    fe v3, v7, p, t, check, ri;
    Ops::sqr(v3, v); Ops::mul(v3, v3, v);                   // v^3
    Ops::sqr(v7, v3); Ops::mul(v7, v7, v);                  // v^7
    fe_pow22523_ops<Ops>(p, v7);
    Ops::mul(t, p, v3);                                     // v^3 (v^7)^((p-5)/8)
    Ops::sqr(check, t); Ops::mul(check, check, v);          // v r^2
    const fe one = fe_one(), i = fe_sqrt_m1();
    fe m1, mi;
    fe_neg(m1, one); fe_neg(mi, i);
    bool correct = fe_eq(check, one);
    bool flipped = fe_eq(check, m1);
    bool flipped_i = fe_eq(check, mi);
    Ops::mul(ri, t, i);
    fe_select(t, t, ri, flipped | flipped_i);
    fe_abs(r, t);
    return correct | flipped;
}
ZK_HD inline bool fe_sqrt_ratio_m1(fe& r, const fe& u, const fe& v) {
    sqrt_ratio_state st; fe p;
    fe_sqrt_ratio_pre(st, u, v);
    fe_pow22523(p, st.t);
    return fe_sqrt_ratio_post(r, st, p, u, v);
}

}  // namespace zk
