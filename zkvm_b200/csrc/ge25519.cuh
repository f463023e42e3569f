// Edwards25519 group law (a = -1, extended coordinates) and the ristretto255 encoding.
//
// Behavioural spec: RFC 9496 section 4.3 (Decode 4.3.1, Encode 4.3.2, element derivation
// 4.3.4) and the Hisil-Wong-Carter-Dawson extended-coordinate formulas ("add-2008-hwcd-3",
// "dbl-2008-hwcd", a=-1).  The formulas are complete on this curve (a square, d non-square),
// so identity and doubling inputs need no special cases -- the same property dalek's
// EdwardsPoint arithmetic relies on (upstream source not mounted).
#pragma once
#include "fe25519.cuh"

namespace zk {

// Extended point (X:Y:Z:T), x = X/Z, y = Y/Z, xy = T/Z.  128 bytes.
struct ge_ext { fe X, Y, Z, T; };
// Affine "Niels" cache entry (y+x, y-x, 2d*x*y) of a point with Z = 1.  96 bytes.
// This is the device-resident point-table format: a mixed add against it costs 7 multiplies.
struct ge_niels { fe yp, ym, t2d; };

ZK_HD ZK_INLINE void ge_identity(ge_ext& r) {
    r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero();
}
ZK_HD ZK_INLINE void ge_niels_identity(ge_niels& r) {
    r.yp = fe_one(); r.ym = fe_one(); r.t2d = fe_zero();
}

// r = p + q (q affine Niels), or p - q when `neg`.  7M.
ZK_HD ZK_INLINE void ge_madd(ge_ext& r, const ge_ext& p, const ge_niels& q, bool neg) {
    fe a, b, c, d, e, f, g, h, qp, qm;
    fe_select(qp, q.yp, q.ym, neg);
    fe_select(qm, q.ym, q.yp, neg);
    fe_sub(a, p.Y, p.X); fe_mul(a, a, qm);
    fe_add(b, p.Y, p.X); fe_mul(b, b, qp);
    fe_mul(c, p.T, q.t2d);
    fe_dbl(d, p.Z);
    fe_sub(e, b, a);
    fe_add(h, b, a);
    fe f0, g0;
    fe_sub(f0, d, c); fe_add(g0, d, c);
    fe_select(f, f0, g0, neg);
    fe_select(g, g0, f0, neg);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.T, e, h); fe_mul(r.Z, f, g);
}

// r = p + q, both extended.  9M.
ZK_HD ZK_INLINE void ge_add(ge_ext& r, const ge_ext& p, const ge_ext& q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(a, p.Y, p.X); fe_sub(t, q.Y, q.X); fe_mul(a, a, t);
    fe_add(b, p.Y, p.X); fe_add(t, q.Y, q.X); fe_mul(b, b, t);
    fe k = fe_d2();
    fe_mul(c, p.T, q.T); fe_mul(c, c, k);
    fe_mul(d, p.Z, q.Z); fe_dbl(d, d);
    fe_sub(e, b, a); fe_sub(f, d, c); fe_add(g, d, c); fe_add(h, b, a);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.T, e, h); fe_mul(r.Z, f, g);
}

// r = 2p.  4S + 4M.
ZK_HD ZK_INLINE void ge_dbl(ge_ext& r, const ge_ext& p) {
    fe a, b, c, e, f, g, h, t;
    fe_sqr(a, p.X); fe_sqr(b, p.Y);
    fe_sqr(c, p.Z); fe_dbl(c, c);
    fe_add(t, p.X, p.Y); fe_sqr(t, t);
    fe_add(h, a, b);            // -H of the a=-1 formula: H = D - B = -(A+B)
    fe_sub(e, t, h);            // E = (X+Y)^2 - A - B
    fe_sub(g, b, a);            // G = D + B = B - A
    fe_sub(f, g, c);            // F = G - C
    fe_neg(h, h);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.T, e, h); fe_mul(r.Z, f, g);
}

ZK_HD ZK_INLINE void ge_neg(ge_ext& r, const ge_ext& p) {
    fe_neg(r.X, p.X); r.Y = p.Y; r.Z = p.Z; fe_neg(r.T, p.T);
}

// Niels form of an affine point (Z == 1 required).
ZK_HD ZK_INLINE void ge_to_niels_affine(ge_niels& r, const fe& x, const fe& y, const fe& t) {
    fe k = fe_d2();
    fe_add(r.yp, y, x); fe_sub(r.ym, y, x); fe_mul(r.t2d, t, k);
}

// Extended form of +-Q from its Niels entry, at the cost of ONE multiply: (X:Y:Z:T) = (2x : 2y : 2 : 2xy), where
// 2xy = t2d / d.  Used to start a bucket accumulator from its first point instead of adding it to the identity.
ZK_HD ZK_INLINE void ge_from_niels(ge_ext& r, const ge_niels& q, bool neg) {
    const fe dinv = {{0xcdc9f843u, 0x25e0f276u, 0x4279542eu, 0x0b5dd698u, 0xcdb9cf66u, 0x2b162114u, 0x14d5ce43u, 0x40907ed2u}};  // 1/d
    fe x, t;
    fe_sub(x, q.yp, q.ym);
    fe_add(r.Y, q.yp, q.ym);
    r.Z = fe_zero(); r.Z.v[0] = 2;
    fe_mul(t, q.t2d, dinv);
    fe_cneg(r.X, x, neg);
    fe_cneg(r.T, t, neg);
}

// ---- ristretto255 ----

// RFC 9496 4.3.1 Decode, split around the inverse square root.  `w` = the 32 bytes as 8 little-endian words.
struct decode_state { fe s, u1, u2, v, arg; bool early_ok; };
ZK_HD ZK_INLINE void ristretto_decode_pre(decode_state& d, const uint32_t w[8]) {
    bool canonical = fe_from_words(d.s, w);
    bool s_neg = (w[0] & 1u) != 0;
    d.early_ok = canonical & !s_neg;
    fe ss, u2s, tmp, one = fe_one(), dd = fe_d();
    fe_sqr(ss, d.s);
    fe_sub(d.u1, one, ss);
    fe_add(d.u2, one, ss);
    fe_sqr(u2s, d.u2);
    fe_sqr(tmp, d.u1); fe_mul(tmp, tmp, dd); fe_neg(tmp, tmp);
    fe_sub(d.v, tmp, u2s);                        // v = -(d u1^2) - u2^2
    fe_mul(d.arg, d.v, u2s);                      // SQRT_RATIO_M1(1, v * u2^2)
}
// On success writes the affine representative (x, y, t = xy), Z = 1.
ZK_HD ZK_INLINE bool ristretto_decode_post(fe& x, fe& y, fe& t, const decode_state& d, const fe& isr, bool was_square) {
    fe dx, dy, tmp;
    fe_mul(dx, isr, d.u2);
    fe_mul(dy, isr, dx); fe_mul(dy, dy, d.v);
    fe_mul(tmp, d.s, dx); fe_dbl(tmp, tmp);
    fe_abs(x, tmp);
    fe_mul(y, d.u1, dy);
    fe_mul(t, x, y);
    bool bad = !d.early_ok | !was_square | fe_is_negative(t) | fe_is_zero(y);
    return !bad;
}
// Returns false on any reject rule.
ZK_HD inline bool ristretto_decode(fe& x, fe& y, fe& t, const uint32_t w[8]) {
    decode_state d; fe isr;
    ristretto_decode_pre(d, w);
    bool was_square = fe_invsqrt(isr, d.arg);
    return ristretto_decode_post(x, y, t, d, isr, was_square);
}
// Two decodes sharing one interleaved exponentiation chain (see fe2).
ZK_HD inline void ristretto_decode_x2(fe& x0, fe& y0, fe& t0, bool& ok0, const uint32_t w0[8],
                                      fe& x1, fe& y1, fe& t1, bool& ok1, const uint32_t w1[8]) {
    decode_state d0, d1; sqrt_ratio_state s0, s1; fe one = fe_one(), isr0, isr1;
    ristretto_decode_pre(d0, w0); ristretto_decode_pre(d1, w1);
    fe_sqrt_ratio_pre(s0, one, d0.arg); fe_sqrt_ratio_pre(s1, one, d1.arg);
    fe2 z, p; z.a = s0.t; z.b = s1.t;
    fe_pow22523_t<fe2>(p, z);
    bool sq0 = fe_sqrt_ratio_post(isr0, s0, p.a, one, d0.arg);
    bool sq1 = fe_sqrt_ratio_post(isr1, s1, p.b, one, d1.arg);
    ok0 = ristretto_decode_post(x0, y0, t0, d0, isr0, sq0);
    ok1 = ristretto_decode_post(x1, y1, t1, d1, isr1, sq1);
}

// RFC 9496 4.3.2 Encode.  Output: canonical 8 little-endian words.  Ops = operation policy (fe_ops_inline in the
// throughput kernels, fe_ops_call in the single-warp tail).
template <class Ops>
ZK_HD inline void ristretto_encode_ops(uint32_t out[8], const ge_ext& p) {
    fe u1, u2, t0, t1, isr, den1, den2, zinv, ix0, iy0, ench, x, y, dinv;
    fe_add(t0, p.Z, p.Y); fe_sub(t1, p.Z, p.Y); Ops::mul(u1, t0, t1);
    Ops::mul(u2, p.X, p.Y);
    Ops::sqr(t0, u2); Ops::mul(t0, t0, u1);
    fe_invsqrt<Ops>(isr, t0);
    Ops::mul(den1, isr, u1);
    Ops::mul(den2, isr, u2);
    Ops::mul(zinv, den1, den2); Ops::mul(zinv, zinv, p.T);
    fe i = fe_sqrt_m1();
    Ops::mul(ix0, p.X, i); Ops::mul(iy0, p.Y, i);
    fe k = fe_invsqrt_a_minus_d();
    Ops::mul(ench, den1, k);
    Ops::mul(t0, p.T, zinv);
    bool rotate = fe_is_negative(t0);
    fe_select(x, p.X, iy0, rotate);
    fe_select(y, p.Y, ix0, rotate);
    fe_select(dinv, den2, ench, rotate);
    Ops::mul(t0, x, zinv);
    fe_cneg(y, y, fe_is_negative(t0));
    fe_sub(t0, p.Z, y); Ops::mul(t0, t0, dinv);
    fe_abs(t0, t0);
    fe_freeze(t0, t0);
    for (int j = 0; j < 8; j++) out[j] = t0.v[j];
}
ZK_HD inline void ristretto_encode(uint32_t out[8], const ge_ext& p) { ristretto_encode_ops<fe_ops_inline>(out, p); }

// RFC 9496 4.3.4 MAP (Elligator 2 on the Jacobi quartic), t already reduced.
ZK_HD inline void ristretto_map(ge_ext& r, const fe& t) {
    fe rr, u, v, s, sp, c, N, w0, w1, w2, w3, tmp, one = fe_one(), dd = fe_d();
    fe i = fe_sqrt_m1();
    fe_sqr(rr, t); fe_mul(rr, rr, i);                    // r = i t^2
    fe_add(u, rr, one); fe k1 = fe_one_minus_d_sq(); fe_mul(u, u, k1);
    fe_mul(tmp, rr, dd); fe_add(tmp, tmp, one); fe_neg(tmp, tmp);   // -1 - r d
    fe_add(v, rr, dd); fe_mul(v, tmp, v);
    bool was_square = fe_sqrt_ratio_m1(s, u, v);
    fe_mul(sp, s, t); fe_abs(sp, sp); fe_neg(sp, sp);    // s' = -|s t|
    fe_select(s, sp, s, was_square);
    fe m1; fe_neg(m1, one);
    fe_select(c, rr, m1, was_square);
    fe_sub(tmp, rr, one); fe_mul(N, c, tmp); fe k2 = fe_d_minus_one_sq(); fe_mul(N, N, k2); fe_sub(N, N, v);
    fe_mul(w0, s, v); fe_dbl(w0, w0);
    fe k3 = fe_sqrt_ad_minus_one(); fe_mul(w1, N, k3);
    fe_sqr(tmp, s);
    fe_sub(w2, one, tmp);
    fe_add(w3, one, tmp);
    fe_mul(r.X, w0, w3); fe_mul(r.Y, w2, w1); fe_mul(r.Z, w1, w3); fe_mul(r.T, w0, w2);
}

// RFC 9496 4.3.4: 64 uniform bytes -> element.  `w` = 16 little-endian words.
ZK_HD inline void ristretto_from_uniform(ge_ext& r, const uint32_t w[16]) {
    fe t1, t2; ge_ext p1, p2;
    fe_from_words(t1, w);        // masks bit 255; loose value < 2^255 is a valid representative mod p
    fe_from_words(t2, w + 8);
    ristretto_map(p1, t1);
    ristretto_map(p2, t2);
    ge_add(r, p1, p2);
}

// Ristretto equality of two extended points (RFC 9496 4.3.3): X1 Y2 == Y1 X2  or  Y1 Y2 == X1 X2.
ZK_HD inline bool ristretto_eq(const ge_ext& p, const ge_ext& q) {
    fe a, b, c, d;
    fe_mul(a, p.X, q.Y); fe_mul(b, p.Y, q.X);
    fe_mul(c, p.Y, q.Y); fe_mul(d, p.X, q.X);
    return fe_eq(a, b) | fe_eq(c, d);
}

}  // namespace zk
