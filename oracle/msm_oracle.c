/* TEST INFRASTRUCTURE ONLY -- CPU oracle.  Nothing under zkvm_b200/ may call, link or load this.
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
 *
 * What it is: a plain-C restatement of the algorithm ZkVM's hot path runs on the CPU -- the
 * variable-time Ristretto255 multiscalar multiplication of curve25519-dalek -- written from the
 * published algorithms because the reference tree is empty:
 *   /root/reference/README.md:1-7 is a "repository has moved" notice, license.txt:1-201 is
 *   Apache-2.0.  curve25519-dalek is a third-party dependency of
 *   interstellar/slingshot; its pinned version is UNKNOWN (no Cargo.lock is mounted), there is no
 *   Rust toolchain and no network, so it can be neither cited by file:line nor compiled.
 *
 * PARITY STATUS: **unpinned against the reference** (no golden vector, test or fixture of the
 * reference exists to check against).  Pinned instead against
 *   - RFC 9496 Appendix A vectors           (tests/golden/rfc9496_vectors.json)
 *   - libsodium 1.0.20 run in this container (tests/golden/libsodium_vectors.json)
 *   - the independent big-integer restatement oracle/ristretto255_ref.py
 * The MSM output is the canonical encoding of a group element (RFC 9496 4.3.2), so it does not
 * depend on which correct algorithm computed it.
 *
 * Algorithm shape (recalled from public knowledge of dalek 1.x-4.x, SURVEY.md Appendix A.1 --
 * "unverified recall", used only so that the CPU baseline has the same cost class):
 *   n < 190  : Straus with width-5 NAFs and 8-entry odd-multiple tables;
 *   otherwise: Pippenger, signed radix-2^w digits, w = 6 (n<500), 7 (n<800), 8 above; per digit
 *              column fill 2^(w-1) buckets, running-sum them, Horner over columns by w doublings.
 * Field arithmetic: radix-2^51, five 64-bit limbs, 128-bit products (dalek's `u64` serial
 * backend shape).  This serial code is THE checker: every parity gate and test compares against it.
 * For the TIMING arm only there is also a vector path in the shape of dalek's AVX2/IFMA backends
 * (four coordinates of a point in the four lanes of a vector, AVX-512 IFMA, bucket accumulation
 * only; see "optional 4-lane IFMA bucket accumulation" below), selected with oracle_set_vector()
 * where the CPU has IFMA and checked against the serial code in tests/test_oracle.py.  dalek's
 * own backends cannot be built here (no Rust); DESIGN.md states this next to every CPU number.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static const fe FE_ZERO = {{0, 0, 0, 0, 0}};
static const fe FE_ONE = {{1, 0, 0, 0, 0}};
static fe FE_D, FE_D2, FE_SQRT_M1, FE_INVSQRT_A_MINUS_D, FE_ONE_MINUS_D_SQ, FE_D_MINUS_ONE_SQ, FE_SQRT_AD_MINUS_ONE;

static void fe_frombytes(fe* h, const uint8_t s[32]) {           /* ignores bit 255 */
    uint64_t w[4];
    memcpy(w, s, 32);
    h->v[0] = w[0] & M51;
    h->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    h->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    h->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    h->v[4] = (w[3] >> 12) & M51;
}
static void fe_carry(fe* h) {
    uint64_t c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += 19 * c;
}
static void fe_tobytes(uint8_t s[32], const fe* f) {
    fe t = *f;
    fe_carry(&t); fe_carry(&t);
    /* t < 2^255 + small; subtract p if t >= p */
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
    t.v[4] &= M51;
    uint64_t w[4];
    w[0] = t.v[0] | (t.v[1] << 51);
    w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
    w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
    w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
    memcpy(s, w, 32);
}
static inline void fe_add(fe* h, const fe* f, const fe* g) { for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i]; }
static inline void fe_sub(fe* h, const fe* f, const fe* g) {     /* f + 16p - g, then weak carry */
    h->v[0] = f->v[0] + 0x7ffffffffffed0ULL - g->v[0];
    h->v[1] = f->v[1] + 0x7ffffffffffff0ULL - g->v[1];
    h->v[2] = f->v[2] + 0x7ffffffffffff0ULL - g->v[2];
    h->v[3] = f->v[3] + 0x7ffffffffffff0ULL - g->v[3];
    h->v[4] = f->v[4] + 0x7ffffffffffff0ULL - g->v[4];
    fe_carry(h);
}
static inline void fe_neg(fe* h, const fe* f) { fe_sub(h, &FE_ZERO, f); }
static void fe_mul(fe* h, const fe* f, const fe* g) {
    uint64_t f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    uint64_t g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
    uint64_t g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
    u128 r0 = (u128)f0 * g0 + (u128)f1 * g4_19 + (u128)f2 * g3_19 + (u128)f3 * g2_19 + (u128)f4 * g1_19;
    u128 r1 = (u128)f0 * g1 + (u128)f1 * g0 + (u128)f2 * g4_19 + (u128)f3 * g3_19 + (u128)f4 * g2_19;
    u128 r2 = (u128)f0 * g2 + (u128)f1 * g1 + (u128)f2 * g0 + (u128)f3 * g4_19 + (u128)f4 * g3_19;
    u128 r3 = (u128)f0 * g3 + (u128)f1 * g2 + (u128)f2 * g1 + (u128)f3 * g0 + (u128)f4 * g4_19;
    u128 r4 = (u128)f0 * g4 + (u128)f1 * g3 + (u128)f2 * g2 + (u128)f3 * g1 + (u128)f4 * g0;
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & M51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & M51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & M51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & M51;
    c = (uint64_t)(r4 >> 51); h->v[4] = (uint64_t)r4 & M51;
    h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static void fe_sq(fe* h, const fe* f) {
    uint64_t f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    uint64_t f0_2 = 2 * f0, f1_2 = 2 * f1, f3_19 = 19 * f3, f4_19 = 19 * f4;
    u128 r0 = (u128)f0 * f0 + (u128)f1_2 * f4_19 + (u128)(2 * f2) * f3_19;
    u128 r1 = (u128)f0_2 * f1 + (u128)(2 * f2) * f4_19 + (u128)f3 * f3_19;
    u128 r2 = (u128)f0_2 * f2 + (u128)f1 * f1 + (u128)(2 * f3) * f4_19;
    u128 r3 = (u128)f0_2 * f3 + (u128)f1_2 * f2 + (u128)f4 * f4_19;
    u128 r4 = (u128)f0_2 * f4 + (u128)f1_2 * f3 + (u128)f2 * f2;
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & M51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & M51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & M51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & M51;
    c = (uint64_t)(r4 >> 51); h->v[4] = (uint64_t)r4 & M51;
    h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static void fe_sqn(fe* h, const fe* f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }
static int fe_isnegative(const fe* f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static int fe_iszero(const fe* f) {
    uint8_t s[32]; fe_tobytes(s, f);
    uint8_t o = 0; for (int i = 0; i < 32; i++) o |= s[i];
    return o == 0;
}
static int fe_eq(const fe* a, const fe* b) { fe t; fe_sub(&t, a, b); return fe_iszero(&t); }
static void fe_abs(fe* h, const fe* f) { if (fe_isnegative(f)) fe_neg(h, f); else *h = *f; }
static void fe_pow22523(fe* r, const fe* z) {                    /* z^(2^252-3) */
    fe t0, t1, t2;
    fe_sq(&t0, z); fe_sqn(&t1, &t0, 2); fe_mul(&t1, z, &t1); fe_mul(&t0, &t0, &t1);
    fe_sq(&t0, &t0); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 5); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 10); fe_mul(&t1, &t1, &t0);
    fe_sqn(&t2, &t1, 20); fe_mul(&t1, &t2, &t1);
    fe_sqn(&t1, &t1, 10); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 50); fe_mul(&t1, &t1, &t0);
    fe_sqn(&t2, &t1, 100); fe_mul(&t1, &t2, &t1);
    fe_sqn(&t1, &t1, 50); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t0, &t0, 2); fe_mul(r, &t0, z);
}
/* RFC 9496 4.2 */
static int fe_sqrt_ratio_m1(fe* r, const fe* u, const fe* v) {
    fe v3, v7, t, check, nu, nui;
    fe_sq(&v3, v); fe_mul(&v3, &v3, v);
    fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);
    fe_mul(&t, u, &v7); fe_pow22523(&t, &t); fe_mul(&t, &t, &v3); fe_mul(&t, &t, u);
    fe_sq(&check, &t); fe_mul(&check, &check, v);
    fe_neg(&nu, u); fe_mul(&nui, &nu, &FE_SQRT_M1);
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &nu), flipped_i = fe_eq(&check, &nui);
    if (flipped || flipped_i) fe_mul(&t, &t, &FE_SQRT_M1);
    fe_abs(r, &t);
    return correct | flipped;
}

/* ---- group ---- */
typedef struct { fe X, Y, Z, T; } ge;                /* extended */
typedef struct { fe YpX, YmX, Z, T2d; } ge_cached;   /* "ProjectiveNiels" */

static void ge_identity(ge* r) { r->X = FE_ZERO; r->Y = FE_ONE; r->Z = FE_ONE; r->T = FE_ZERO; }
static void ge_to_cached(ge_cached* c, const ge* p) {
    fe_add(&c->YpX, &p->Y, &p->X); fe_sub(&c->YmX, &p->Y, &p->X); c->Z = p->Z; fe_mul(&c->T2d, &p->T, &FE_D2);
}
static void ge_add_cached(ge* r, const ge* p, const ge_cached* q, int neg) {   /* 8M */
    fe a, b, c, d, e, f, g, h;
    fe_sub(&a, &p->Y, &p->X); fe_add(&b, &p->Y, &p->X);
    if (!neg) { fe_mul(&a, &a, &q->YmX); fe_mul(&b, &b, &q->YpX); }
    else      { fe_mul(&a, &a, &q->YpX); fe_mul(&b, &b, &q->YmX); }
    fe_mul(&c, &p->T, &q->T2d);
    fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d); fe_carry(&d);
    fe_sub(&e, &b, &a); fe_add(&h, &b, &a);
    if (!neg) { fe_sub(&f, &d, &c); fe_add(&g, &d, &c); } else { fe_add(&f, &d, &c); fe_sub(&g, &d, &c); }
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}
static void ge_add(ge* r, const ge* p, const ge* q) { ge_cached c; ge_to_cached(&c, q); ge_add_cached(r, p, &c, 0); }
static void ge_dbl(ge* r, const ge* p) {                                        /* 4S + 4M */
    fe a, b, c, e, f, g, h, t;
    fe_sq(&a, &p->X); fe_sq(&b, &p->Y); fe_sq(&c, &p->Z); fe_add(&c, &c, &c);
    fe_add(&t, &p->X, &p->Y); fe_sq(&t, &t);
    fe_add(&h, &a, &b); fe_sub(&e, &t, &h); fe_sub(&g, &b, &a); fe_sub(&f, &g, &c); fe_neg(&h, &h);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}

/* RFC 9496 4.3.1; returns 1 on success */
static int rist_decode(ge* p, const uint8_t in[32]) {
    fe s; uint8_t chk[32];
    fe_frombytes(&s, in);
    fe_tobytes(chk, &s);
    int canonical = memcmp(chk, in, 32) == 0;       /* also rejects bit 255 */
    int s_neg = in[0] & 1;
    fe ss, u1, u2, u2s, v, isr, dx, dy, tmp;
    fe_sq(&ss, &s); fe_sub(&u1, &FE_ONE, &ss); fe_add(&u2, &FE_ONE, &ss); fe_sq(&u2s, &u2);
    fe_sq(&tmp, &u1); fe_mul(&tmp, &tmp, &FE_D); fe_neg(&tmp, &tmp); fe_sub(&v, &tmp, &u2s);
    fe_mul(&tmp, &v, &u2s);
    int was_square = fe_sqrt_ratio_m1(&isr, &FE_ONE, &tmp);
    fe_mul(&dx, &isr, &u2); fe_mul(&dy, &isr, &dx); fe_mul(&dy, &dy, &v);
    fe_mul(&tmp, &s, &dx); fe_add(&tmp, &tmp, &tmp); fe_abs(&p->X, &tmp);
    fe_mul(&p->Y, &u1, &dy); p->Z = FE_ONE; fe_mul(&p->T, &p->X, &p->Y);
    return canonical && !s_neg && was_square && !fe_isnegative(&p->T) && !fe_iszero(&p->Y);
}
/* RFC 9496 4.3.2 */
static void rist_encode(uint8_t out[32], const ge* p) {
    fe u1, u2, t0, t1, isr, den1, den2, zinv, ix0, iy0, ench, x, y, dinv;
    fe_add(&t0, &p->Z, &p->Y); fe_sub(&t1, &p->Z, &p->Y); fe_mul(&u1, &t0, &t1);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&t0, &u2); fe_mul(&t0, &t0, &u1);
    fe_sqrt_ratio_m1(&isr, &FE_ONE, &t0);
    fe_mul(&den1, &isr, &u1); fe_mul(&den2, &isr, &u2);
    fe_mul(&zinv, &den1, &den2); fe_mul(&zinv, &zinv, &p->T);
    fe_mul(&ix0, &p->X, &FE_SQRT_M1); fe_mul(&iy0, &p->Y, &FE_SQRT_M1);
    fe_mul(&ench, &den1, &FE_INVSQRT_A_MINUS_D);
    fe_mul(&t0, &p->T, &zinv);
    if (fe_isnegative(&t0)) { x = iy0; y = ix0; dinv = ench; } else { x = p->X; y = p->Y; dinv = den2; }
    fe_mul(&t0, &x, &zinv);
    if (fe_isnegative(&t0)) fe_neg(&y, &y);
    fe_sub(&t0, &p->Z, &y); fe_mul(&t0, &t0, &dinv); fe_abs(&t0, &t0);
    fe_tobytes(out, &t0);
}
/* RFC 9496 4.3.4 MAP */
static void rist_map(ge* r, const fe* t) {
    fe rr, u, v, s, sp, c, N, w0, w1, w2, w3, tmp;
    fe_sq(&rr, t); fe_mul(&rr, &rr, &FE_SQRT_M1);
    fe_add(&u, &rr, &FE_ONE); fe_mul(&u, &u, &FE_ONE_MINUS_D_SQ);
    fe_mul(&tmp, &rr, &FE_D); fe_add(&tmp, &tmp, &FE_ONE); fe_neg(&tmp, &tmp);
    fe_add(&v, &rr, &FE_D); fe_mul(&v, &tmp, &v);
    int was_square = fe_sqrt_ratio_m1(&s, &u, &v);
    fe_mul(&sp, &s, t); fe_abs(&sp, &sp); fe_neg(&sp, &sp);
    if (!was_square) { s = sp; c = rr; } else { fe_neg(&c, &FE_ONE); }
    fe_sub(&tmp, &rr, &FE_ONE); fe_mul(&N, &c, &tmp); fe_mul(&N, &N, &FE_D_MINUS_ONE_SQ); fe_sub(&N, &N, &v);
    fe_mul(&w0, &s, &v); fe_add(&w0, &w0, &w0);
    fe_mul(&w1, &N, &FE_SQRT_AD_MINUS_ONE);
    fe_sq(&tmp, &s); fe_sub(&w2, &FE_ONE, &tmp); fe_add(&w3, &FE_ONE, &tmp);
    fe_mul(&r->X, &w0, &w3); fe_mul(&r->Y, &w2, &w1); fe_mul(&r->Z, &w1, &w3); fe_mul(&r->T, &w0, &w2);
}

/* ---- scalars ---- */
/* s (256-bit LE) -> s mod l, as 4 x u64.  l = 2^252 + DL. */
static const uint64_t DL[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL};
static void sc_reduce256(uint64_t r[4], const uint8_t s[32]) {
    uint64_t w[4]; memcpy(w, s, 32);
    uint64_t q = w[3] >> 60; w[3] &= 0x0fffffffffffffffULL;
    /* t = q * DL (3 words) */
    u128 m0 = (u128)DL[0] * q, m1 = (u128)DL[1] * q + (uint64_t)(m0 >> 64);
    uint64_t t[4] = {(uint64_t)m0, (uint64_t)m1, (uint64_t)(m1 >> 64), 0};
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)w[i] - t[i] - br;
        r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1;
    }
    if (br) {   /* add l back */
        uint64_t l[4] = {DL[0], DL[1], 0, 0x1000000000000000ULL};
        u128 c = 0;
        for (int i = 0; i < 4; i++) { c += (u128)r[i] + l[i]; r[i] = (uint64_t)c; c >>= 64; }
    }
}
/* signed radix-2^w digits of a reduced scalar; returns digit count */
static int sc_radix_2w(int8_t* digits, const uint64_t s[4], int w) {
    int count = (256 + w - 1) / w + 1;
    int carry = 0;
    for (int i = 0; i < count; i++) {
        int bit = i * w, word = bit >> 6, sh = bit & 63;
        uint64_t raw = 0;
        if (word < 4) { raw = s[word] >> sh; if (sh && word + 1 < 4) raw |= s[word + 1] << (64 - sh); }
        int coef = (int)(raw & ((1u << w) - 1)) + carry;
        carry = (coef + (1 << (w - 1))) >> w;          /* coef > 2^(w-1)-ish rounds up */
        digits[i] = (int8_t)(coef - (carry << w));
    }
    return count;
}
/* width-5 non-adjacent form, 256 entries */
static void sc_naf5(int8_t naf[257], const uint64_t s[4]) {
    memset(naf, 0, 257);
    uint64_t x[5] = {s[0], s[1], s[2], s[3], 0};
    int pos = 0, carry = 0;
    while (pos < 257) {
        int word = pos >> 6, sh = pos & 63;
        uint64_t bits = x[word] >> sh;
        if (sh > 59 && word + 1 < 5) bits |= x[word + 1] << (64 - sh);
        int window = carry + (int)(bits & 31);
        if ((window & 1) == 0) { pos += 1; continue; }
        if (window < 16) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)(window - 32); }
        pos += 5;
    }
}

/* ---- MSM over decompressed points ---- */
static void msm_straus(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
    int8_t (*nafs)[257] = malloc(n * 257);
    ge_cached (*tabs)[8] = malloc(n * sizeof(ge_cached[8]));
    for (size_t i = 0; i < n; i++) {
        uint64_t s[4]; sc_reduce256(s, scalars + 32 * i); sc_naf5(nafs[i], s);
        ge p2, cur = pts[i]; ge_dbl(&p2, &pts[i]);
        ge_to_cached(&tabs[i][0], &cur);
        for (int k = 1; k < 8; k++) { ge_add(&cur, &cur, &p2); ge_to_cached(&tabs[i][k], &cur); }
    }
    ge r; ge_identity(&r);
    int top = 256;
    for (; top >= 0; top--) { int any = 0; for (size_t i = 0; i < n && !any; i++) any = nafs[i][top] != 0; if (any) break; }
    for (int b = top; b >= 0; b--) {
        ge_dbl(&r, &r);
        for (size_t i = 0; i < n; i++) {
            int d = nafs[i][b];
            if (d > 0) ge_add_cached(&r, &r, &tabs[i][d >> 1], 0);
            else if (d < 0) ge_add_cached(&r, &r, &tabs[i][(-d) >> 1], 1);
        }
    }
    *out = r;
    free(nafs); free(tabs);
}
static void msm_pippenger(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
    int w = n < 500 ? 6 : n < 800 ? 7 : 8;
    int nb = 1 << (w - 1);
    int max_digits = (256 + w - 1) / w + 1;
    int8_t* digits = malloc(n * (size_t)max_digits);
    ge_cached* cached = malloc(n * sizeof(ge_cached));
    ge* buckets = malloc(nb * sizeof(ge));
    int count = max_digits;
    for (size_t i = 0; i < n; i++) {
        uint64_t s[4]; sc_reduce256(s, scalars + 32 * i);
        count = sc_radix_2w(digits + i * max_digits, s, w);
        ge_to_cached(&cached[i], &pts[i]);
    }
    ge total; ge_identity(&total);
    for (int col = count - 1; col >= 0; col--) {
        for (int b = 0; b < nb; b++) ge_identity(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            int d = digits[i * max_digits + col];
            if (d > 0) ge_add_cached(&buckets[d - 1], &buckets[d - 1], &cached[i], 0);
            else if (d < 0) ge_add_cached(&buckets[-d - 1], &buckets[-d - 1], &cached[i], 1);
        }
        ge run = buckets[nb - 1], sum = buckets[nb - 1];
        for (int b = nb - 2; b >= 0; b--) { ge_add(&run, &run, &buckets[b]); ge_add(&sum, &sum, &run); }
        for (int k = 0; k < w; k++) ge_dbl(&total, &total);
        ge_add(&total, &total, &sum);
    }
    *out = total;
    free(digits); free(cached); free(buckets);
}
/* ---- optional 4-lane IFMA bucket accumulation -------------------------------------------------------------------
 * dalek ships vector backends (AVX2, AVX-512 IFMA) that hold the four coordinates of a point in the four 64-bit lanes
 * of a vector and run the HWCD addition as two 4-way field multiplications plus lane shuffles.  This is that shape for
 * the one loop that dominates a Pippenger MSM -- bucket[digit] += cached point -- on AVX-512 IFMA (vpmadd52luq/huq,
 * radix 2^51, five limbs per lane); bucket reduction and the Horner over columns stay scalar (2^w additions and w
 * doublings per column against n).  Selected at run time (oracle_set_vector) and only where the CPU has it; checked
 * against the scalar path in tests/test_oracle.py.  It exists so that the CPU baseline can also be quoted for the
 * reference's vectorised build, not only for its default serial one ("AVX2/IFMA backend as built", BASELINE.json). */
static int g_use_vec = 0;
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define VEC __attribute__((target("avx2,avx512f,avx512vl,avx512ifma")))
typedef struct { __m256i l[5]; } f4;          /* lanes: X, Y, T, Z -- or, cached: Y-X, Y+X, 2dT, 2Z */
static int vec_available(void) {
    return __builtin_cpu_supports("avx512ifma") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx2");
}
VEC static inline __m256i v_mul19(__m256i v) { return _mm256_add_epi64(_mm256_add_epi64(_mm256_slli_epi64(v, 4), _mm256_slli_epi64(v, 1)), v); }
/* limbs < 2^62 in, limbs < 2^51 + 2^15 out */
VEC static inline void f4_carry(f4* a) {
    const __m256i mask = _mm256_set1_epi64x((long long)M51);
    __m256i c;
    c = _mm256_srli_epi64(a->l[0], 51); a->l[0] = _mm256_and_si256(a->l[0], mask); a->l[1] = _mm256_add_epi64(a->l[1], c);
    c = _mm256_srli_epi64(a->l[1], 51); a->l[1] = _mm256_and_si256(a->l[1], mask); a->l[2] = _mm256_add_epi64(a->l[2], c);
    c = _mm256_srli_epi64(a->l[2], 51); a->l[2] = _mm256_and_si256(a->l[2], mask); a->l[3] = _mm256_add_epi64(a->l[3], c);
    c = _mm256_srli_epi64(a->l[3], 51); a->l[3] = _mm256_and_si256(a->l[3], mask); a->l[4] = _mm256_add_epi64(a->l[4], c);
    c = _mm256_srli_epi64(a->l[4], 51); a->l[4] = _mm256_and_si256(a->l[4], mask); a->l[0] = _mm256_add_epi64(a->l[0], v_mul19(c));
}
/* r = x * y lane by lane; limbs of x, y < 2^52.  x_i y_j = lo52 + 2^52 hi52 and 2^52 = 2 * 2^51: the high halves go to
 * the next limb doubled; limbs 5..9 fold back with 2^255 = 19. */
VEC static inline void f4_mul(f4* r, const f4* x, const f4* y) {
    const __m256i zero = _mm256_setzero_si256();
    __m256i lo[9], hi[9], z[10];
    for (int k = 0; k < 9; k++) { lo[k] = zero; hi[k] = zero; }
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) {
            lo[i + j] = _mm256_madd52lo_epu64(lo[i + j], x->l[i], y->l[j]);
            hi[i + j] = _mm256_madd52hi_epu64(hi[i + j], x->l[i], y->l[j]);
        }
    z[0] = lo[0];
    for (int k = 1; k < 9; k++) z[k] = _mm256_add_epi64(lo[k], _mm256_slli_epi64(hi[k - 1], 1));
    z[9] = _mm256_slli_epi64(hi[8], 1);
    for (int k = 0; k < 5; k++) r->l[k] = _mm256_add_epi64(z[k], v_mul19(z[k + 5]));     /* < 2^56 + 19 * 2^56 */
    f4_carry(r);
}
/* 2p, limb by limb: keeps a - b + 2p non-negative for limbs a, b < 2^51 + 2^15 */
VEC static inline __m256i v_bias(int limb) { return _mm256_set1_epi64x(limb == 0 ? 0xfffffffffffdaLL : 0xffffffffffffeLL); }
/* acc += q (cached: Y-X, Y+X, 2dT, 2Z), the unified a = -1 extended addition as two 4-way multiplications */
VEC static inline void f4_add_cached(f4* acc, const f4* q) {
    f4 u, m, t, L, R;
    for (int i = 0; i < 5; i++) {
        __m256i p = acc->l[i], ps = _mm256_permute4x64_epi64(p, 0xE1);                     /* (Y, X, T, Z) */
        __m256i sum = _mm256_add_epi64(p, ps), dif = _mm256_add_epi64(_mm256_sub_epi64(ps, p), v_bias(i));
        u.l[i] = _mm256_blend_epi32(_mm256_blend_epi32(p, dif, 0x03), sum, 0x0C);          /* Y-X | Y+X | T | Z */
    }
    f4_carry(&u);
    f4_mul(&m, &u, q);                                                                       /* A | B | C | D */
    for (int i = 0; i < 5; i++) {
        __m256i v = m.l[i], vs = _mm256_permute4x64_epi64(v, 0xB1);                        /* (B, A, D, C) */
        __m256i sum = _mm256_add_epi64(v, vs), dif = _mm256_add_epi64(_mm256_sub_epi64(vs, v), v_bias(i));
        t.l[i] = _mm256_blend_epi32(sum, dif, 0x33);                                         /* E=B-A | H=A+B | F=D-C | G=C+D */
    }
    f4_carry(&t);
    for (int i = 0; i < 5; i++) {
        L.l[i] = _mm256_permute4x64_epi64(t.l[i], 0x8C);                                     /* E | G | E | F */
        R.l[i] = _mm256_permute4x64_epi64(t.l[i], 0xD6);                                     /* F | H | H | G */
    }
    f4_mul(acc, &L, &R);                                                                     /* X3=EF | Y3=GH | T3=EH | Z3=FG */
}
/* -q: swap Y-X and Y+X, negate 2dT */
VEC static inline void f4_neg_cached(f4* r, const f4* q) {
    for (int i = 0; i < 5; i++) {
        __m256i s = _mm256_permute4x64_epi64(q->l[i], 0xE1);
        r->l[i] = _mm256_blend_epi32(s, _mm256_sub_epi64(v_bias(i), s), 0x30);             /* lane 2 (T) = 2p - t, < 2^52 */
    }
}
VEC static void f4_pack(f4* r, const fe* a, const fe* b, const fe* c, const fe* d) {
    for (int i = 0; i < 5; i++) r->l[i] = _mm256_set_epi64x((long long)d->v[i], (long long)c->v[i], (long long)b->v[i], (long long)a->v[i]);
}
VEC static void msm_pippenger_vec(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
    int w = n < 500 ? 6 : n < 800 ? 7 : 8;
    int nb = 1 << (w - 1);
    int max_digits = (256 + w - 1) / w + 1;
    int8_t* digits = malloc(n * (size_t)max_digits);
    f4* cached = aligned_alloc(32, n * sizeof(f4));
    f4* buckets = aligned_alloc(32, nb * sizeof(f4));
    int count = max_digits;
    for (size_t i = 0; i < n; i++) {
        uint64_t s[4]; sc_reduce256(s, scalars + 32 * i);
        count = sc_radix_2w(digits + i * max_digits, s, w);
        fe ymx, ypx, t2d, z2;
        fe_sub(&ymx, &pts[i].Y, &pts[i].X); fe_add(&ypx, &pts[i].Y, &pts[i].X); fe_carry(&ypx);
        fe_mul(&t2d, &pts[i].T, &FE_D2); fe_add(&z2, &pts[i].Z, &pts[i].Z); fe_carry(&z2);
        f4_pack(&cached[i], &ymx, &ypx, &t2d, &z2);
    }
    f4 ident; f4_pack(&ident, &FE_ZERO, &FE_ONE, &FE_ZERO, &FE_ONE);
    ge total; ge_identity(&total);
    for (int col = count - 1; col >= 0; col--) {
        for (int b = 0; b < nb; b++) buckets[b] = ident;
        for (size_t i = 0; i < n; i++) {
            int d = digits[i * max_digits + col];
            if (d > 0) f4_add_cached(&buckets[d - 1], &cached[i]);
            else if (d < 0) { f4 nq; f4_neg_cached(&nq, &cached[i]); f4_add_cached(&buckets[-d - 1], &nq); }
        }
        /* back to the scalar representation for the running sums */
        ge run, sum;
        for (int b = nb - 1; b >= 0; b--) {
            uint64_t lane[5][4]; ge g;
            for (int i = 0; i < 5; i++) _mm256_storeu_si256((__m256i*)lane[i], buckets[b].l[i]);
            for (int i = 0; i < 5; i++) { g.X.v[i] = lane[i][0]; g.Y.v[i] = lane[i][1]; g.T.v[i] = lane[i][2]; g.Z.v[i] = lane[i][3]; }
            if (b == nb - 1) { run = g; sum = g; } else { ge_add(&run, &run, &g); ge_add(&sum, &sum, &run); }
        }
        for (int k = 0; k < w; k++) ge_dbl(&total, &total);
        ge_add(&total, &total, &sum);
    }
    *out = total;
    free(digits); free(cached); free(buckets);
}
#else
static int vec_available(void) { return 0; }
static void msm_pippenger_vec(ge* out, const uint8_t* scalars, const ge* pts, size_t n) { msm_pippenger(out, scalars, pts, n); }
#endif
/* on != 0: use the 4-lane IFMA bucket accumulation where the CPU has it.  Returns whether it is in use. */
int oracle_set_vector(int on) { g_use_vec = on && vec_available(); return g_use_vec; }

static void msm_single(ge* out, const uint8_t* scalars, const ge* pts, size_t n) {
    if (n == 0) { ge_identity(out); return; }
    if (n < 190) msm_straus(out, scalars, pts, n);
    else if (g_use_vec) msm_pippenger_vec(out, scalars, pts, n);
    else msm_pippenger(out, scalars, pts, n);
}

typedef struct { const uint8_t* scalars; const ge* pts; size_t n; ge out; } msm_job;
static void* msm_worker(void* a) { msm_job* j = a; msm_single(&j->out, j->scalars, j->pts, j->n); return NULL; }

typedef struct { const uint8_t* in; ge* out; size_t n; int bad; size_t bad_index; } dec_job;
static void* dec_worker(void* a) {
    dec_job* j = a; j->bad = 0;
    for (size_t i = 0; i < j->n; i++)
        if (!rist_decode(&j->out[i], j->in + 32 * i) && !j->bad) { j->bad = 1; j->bad_index = i; }
    return NULL;
}

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void init_consts(void) {
    static const uint8_t d[32] = {0xa3,0x78,0x59,0x13,0xca,0x4d,0xeb,0x75,0xab,0xd8,0x41,0x41,0x4d,0x0a,0x70,0x00,0x98,0xe8,0x79,0x77,0x79,0x40,0xc7,0x8c,0x73,0xfe,0x6f,0x2b,0xee,0x6c,0x03,0x52};
    static const uint8_t sm1[32] = {0xb0,0xa0,0x0e,0x4a,0x27,0x1b,0xee,0xc4,0x78,0xe4,0x2f,0xad,0x06,0x18,0x43,0x2f,0xa7,0xd7,0xfb,0x3d,0x99,0x00,0x4d,0x2b,0x0b,0xdf,0xc1,0x4f,0x80,0x24,0x83,0x2b};
    fe_frombytes(&FE_D, d); fe_frombytes(&FE_SQRT_M1, sm1);
    fe_add(&FE_D2, &FE_D, &FE_D); fe_carry(&FE_D2);
    fe t, u;
    fe_sq(&t, &FE_D); fe_sub(&FE_ONE_MINUS_D_SQ, &FE_ONE, &t);
    fe_sub(&t, &FE_D, &FE_ONE); fe_sq(&FE_D_MINUS_ONE_SQ, &t);
    /* invsqrt(a - d) = invsqrt(-1 - d) */
    fe_add(&t, &FE_D, &FE_ONE); fe_neg(&t, &t);
    fe_sqrt_ratio_m1(&FE_INVSQRT_A_MINUS_D, &FE_ONE, &t);
    /* sqrt(a*d - 1) = sqrt(-d - 1); RFC 9496 lists the odd root */
    fe_sqrt_ratio_m1(&u, &t, &FE_ONE);
    fe_neg(&FE_SQRT_AD_MINUS_ONE, &u);
}

/* ---- exported API (ctypes) ---- */
int oracle_decompress(const uint8_t* points32, size_t n, int threads, void* ge_out /* n * sizeof(ge)=160 B */, size_t* bad_index) {
    pthread_once(&g_once, init_consts);
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    dec_job* jobs = calloc(threads, sizeof(dec_job)); pthread_t* th = malloc(threads * sizeof(pthread_t));
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per; if (lo > n) lo = n;
        jobs[t].in = points32 + 32 * lo; jobs[t].out = (ge*)ge_out + lo; jobs[t].n = hi - lo;
        if (t) pthread_create(&th[t], NULL, dec_worker, &jobs[t]);
    }
    dec_worker(&jobs[0]);
    int bad = 0;
    for (int t = 0; t < threads; t++) {
        if (t) pthread_join(th[t], NULL);
        if (jobs[t].bad && !bad) { bad = 1; if (bad_index) *bad_index = (size_t)t * per + jobs[t].bad_index; }
    }
    free(jobs); free(th);
    return bad;
}
size_t oracle_ge_size(void) { return sizeof(ge); }

/* MSM over already-decompressed points (what dalek's vartime_multiscalar_mul takes), index-range sharded over threads */
void oracle_msm_decompressed(const uint8_t* scalars32, const void* ge_in, size_t n, int threads, uint8_t out32[32]) {
    pthread_once(&g_once, init_consts);
    if (threads < 1) threads = 1;
    if ((size_t)threads > n / 256 + 1) threads = (int)(n / 256 + 1);
    msm_job* jobs = calloc(threads, sizeof(msm_job)); pthread_t* th = malloc(threads * sizeof(pthread_t));
    size_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per; if (lo > n) lo = n;
        jobs[t].scalars = scalars32 + 32 * lo; jobs[t].pts = (const ge*)ge_in + lo; jobs[t].n = hi - lo;
        if (t) pthread_create(&th[t], NULL, msm_worker, &jobs[t]);
    }
    msm_worker(&jobs[0]);
    ge acc = jobs[0].out;
    for (int t = 1; t < threads; t++) { pthread_join(th[t], NULL); ge_add(&acc, &acc, &jobs[t].out); }
    rist_encode(out32, &acc);
    free(jobs); free(th);
}
/* decode + MSM + encode; returns 0 ok, 1 if some encoding is invalid (optional_multiscalar_mul -> None) */
int oracle_msm(const uint8_t* scalars32, const uint8_t* points32, size_t n, int threads, uint8_t out32[32]) {
    pthread_once(&g_once, init_consts);
    ge* pts = malloc((n ? n : 1) * sizeof(ge));
    int bad = oracle_decompress(points32, n, threads, pts, NULL);
    if (!bad) oracle_msm_decompressed(scalars32, pts, n, threads, out32); else memset(out32, 0, 32);
    free(pts);
    return bad;
}
void oracle_from_uniform(const uint8_t* in64, size_t n, uint8_t* out32) {
    pthread_once(&g_once, init_consts);
    for (size_t i = 0; i < n; i++) {
        fe t1, t2; ge p1, p2, r;
        fe_frombytes(&t1, in64 + 64 * i); fe_frombytes(&t2, in64 + 64 * i + 32);
        rist_map(&p1, &t1); rist_map(&p2, &t2); ge_add(&r, &p1, &p2);
        rist_encode(out32 + 32 * i, &r);
    }
}
/* out = s * P (double-and-add on the reduced scalar); returns 1 if P is invalid */
int oracle_scalarmult(const uint8_t s32[32], const uint8_t p32[32], uint8_t out32[32]) {
    pthread_once(&g_once, init_consts);
    ge p, r; if (!rist_decode(&p, p32)) return 1;
    uint64_t s[4]; sc_reduce256(s, s32);
    ge_identity(&r);
    for (int b = 255; b >= 0; b--) { ge_dbl(&r, &r); if ((s[b >> 6] >> (b & 63)) & 1) ge_add(&r, &r, &p); }
    rist_encode(out32, &r);
    return 0;
}
/* out = sum of n encoded points; returns 1 if any is invalid */
int oracle_sum(const uint8_t* p32, size_t n, uint8_t out32[32]) {
    pthread_once(&g_once, init_consts);
    ge acc, p; ge_identity(&acc);
    for (size_t i = 0; i < n; i++) { if (!rist_decode(&p, p32 + 32 * i)) return 1; ge_add(&acc, &acc, &p); }
    rist_encode(out32, &acc);
    return 0;
}
int oracle_is_valid(const uint8_t p32[32]) { pthread_once(&g_once, init_consts); ge p; return rist_decode(&p, p32); }
