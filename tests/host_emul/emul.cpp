// TEST HARNESS ONLY.  Compiles the device headers (zkvm_b200/csrc/*.cuh) for the host, where the
// PTX carry chains are replaced by their portable rendering, so the limb schedules and the
// RFC 9496 routines can be checked against the oracle without a GPU.  Never linked into the
// product library.
#include <string.h>
#include "../../zkvm_b200/csrc/ge25519.cuh"
#include "../../zkvm_b200/csrc/recode.cuh"
using namespace zk;

static void ld(fe& r, const uint8_t* b) { memcpy(r.v, b, 32); }
static void st(uint8_t* b, const fe& r) { memcpy(b, r.v, 32); }

extern "C" {
// op: 0 mul, 1 sqr, 2 add, 3 sub, 4 freeze(a), 5 invert(a), 6 neg, 7 pow22523
void emul_fe_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe x, y, r; ld(x, a + 32 * i); ld(y, b + 32 * i);
        switch (op) {
            case 0: fe_mul(r, x, y); break;
            case 1: fe_sqr(r, x); break;
            case 2: fe_add(r, x, y); break;
            case 3: fe_sub(r, x, y); break;
            case 4: fe_freeze(r, x); break;
            case 5: fe_invert(r, x); break;
            case 6: fe_neg(r, x); break;
            case 7: fe_pow22523(r, x); break;
        }
        st(out + 32 * i, r);
    }
}
// returns 1 if ok; writes x,y,t (loose) -- callers compare after reducing mod p
int emul_decode(const uint8_t* in32, uint8_t* xyt96) {
    uint32_t w[8]; memcpy(w, in32, 32);
    fe x, y, t; bool ok = ristretto_decode(x, y, t, w);
    st(xyt96, x); st(xyt96 + 32, y); st(xyt96 + 64, t);
    return ok ? 1 : 0;
}
void emul_encode(const uint8_t* ext128, uint8_t* out32) {
    ge_ext p; ld(p.X, ext128); ld(p.Y, ext128 + 32); ld(p.Z, ext128 + 64); ld(p.T, ext128 + 96);
    uint32_t w[8]; ristretto_encode(w, p); memcpy(out32, w, 32);
}
void emul_from_uniform(const uint8_t* in64, uint8_t* out32) {
    uint32_t w[16]; memcpy(w, in64, 64);
    ge_ext p; ristretto_from_uniform(p, w);
    uint32_t o[8]; ristretto_encode(o, p); memcpy(out32, o, 32);
}
// naive double-and-add MSM through decode -> niels -> madd/dbl -> encode; returns 0 ok, 1 invalid point
int emul_msm(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out32) {
    ge_ext acc; ge_identity(acc);
    for (size_t i = 0; i < n; i++) {
        uint32_t w[8]; memcpy(w, points + 32 * i, 32);
        fe x, y, t; if (!ristretto_decode(x, y, t, w)) return 1;
        ge_niels q; ge_to_niels_affine(q, x, y, t);
        ge_ext r; ge_identity(r);
        const uint8_t* s = scalars + 32 * i;
        for (int bit = 255; bit >= 0; bit--) {
            ge_dbl(r, r);
            if ((s[bit >> 3] >> (bit & 7)) & 1) ge_madd(r, r, q, false);
        }
        // exercise the negated mixed add too: acc += r via  acc = acc - (-r) is not available for ext,
        // so add r, then add q and subtract q again
        ge_add(acc, acc, r);
        ge_madd(acc, acc, q, false); ge_madd(acc, acc, q, true);
        // ge_from_niels: +Q and -Q built from the table entry must cancel
        ge_ext pq, nq; ge_from_niels(pq, q, false); ge_from_niels(nq, q, true);
        ge_add(acc, acc, pq); ge_add(acc, acc, nq);
        ge_add(acc, acc, pq); ge_madd(acc, acc, q, true);
    }
    uint32_t o[8]; ristretto_encode(o, acc); memcpy(out32, o, 32);
    return 0;
}
// the sort kernels' recoding (recode.cuh): scalar32 -> reduced scalar (32 bytes), W signed digits, their offsets and widths
void emul_recode(const uint8_t* scalar32, int W, uint8_t* reduced32, int32_t* digits, int32_t* offs, int32_t* widths) {
    uint32_t s[8]; memcpy(s, scalar32, 32);
    scalar_reduce(s);
    memcpy(reduced32, s, 32);
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        win_geom g = window_geom(W, w);
        digits[w] = next_digit(s, g, carry);
        offs[w] = g.off; widths[w] = g.width;
    }
    digits[W] = (int32_t)carry;      // must be 0: the 254th bit absorbs the top digit's carry
}
}
