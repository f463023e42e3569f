// Scalar recoding of the Pippenger MSM: reduction mod l, balanced window geometry, signed digits.
//
// Shared by the sort kernels (msm.cu) and, compiled for the host, by the CPU tests (tests/host_emul): the digit
// identity  sum_w d_w 2^(off_w) = s mod l  is checked there against big integers for every window count.
// Behavioural spec: the group order l of RFC 9496 section 4 (= RFC 8032's L); nothing here is derived from the
// reference tree, which holds no source for this path.
#pragma once
#include <stdint.h>
#include "fe25519.cuh"      // ZK_HD / ZK_INLINE

namespace zk {

// Window geometry.  A scalar (< 2^253 after reduction) is cut into W signed digits covering 254 bits; the widths are
// balanced -- the first 254 % W windows are one bit wider than the rest -- so no window is a short stub whose points
// pile up in a handful of buckets (with a fixed width c, 253 = 19*13 + 6 leaves a 6-bit top window: 64 buckets
// holding 1/20 of all entries).  The 254th bit is always zero, which absorbs the top digit's carry.
struct win_geom { int off, width; };
ZK_HD ZK_INLINE win_geom window_geom(int W, int w) {
    int base = 254 / W, extra = 254 % W;
    win_geom g; g.width = base + (w < extra ? 1 : 0); g.off = w * base + (w < extra ? w : extra);
    return g;
}
// s (any 256-bit value) -> s mod l, l = 2^252 + delta.  q = floor(s / 2^252) over-estimates the quotient by at most 1.
ZK_HD ZK_INLINE void scalar_reduce(uint32_t s[8]) {
    const uint32_t DL[4] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
    uint32_t q = s[7] >> 28;
    s[7] &= 0x0fffffffu;
    uint32_t t[5]; uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { c += (uint64_t)DL[i] * q; t[i] = (uint32_t)c; c >>= 32; }
    t[4] = (uint32_t)c;
    int64_t br = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { br += (int64_t)s[i] - (i < 5 ? t[i] : 0u); s[i] = (uint32_t)br; br >>= 32; }
    uint32_t m = (uint32_t)br;            // 0 or 0xffffffff: went negative -> add l back
    c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t li = (i < 4 ? DL[i] : (i == 7 ? 0x10000000u : 0u)) & m;
        c += (uint64_t)s[i] + li; s[i] = (uint32_t)c; c >>= 32;
    }
}

// Signed digit of window g of s, given the carry into it.  Digits lie in [-2^(width-1), 2^(width-1)].
ZK_HD ZK_INLINE int next_digit(const uint32_t* s, win_geom g, uint32_t& carry) {
    int word = g.off >> 5, sh = g.off & 31;
    uint32_t lo = word < 8 ? s[word] : 0u, hi = word + 1 < 8 ? s[word + 1] : 0u;
    uint32_t raw = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh) & ((1u << g.width) - 1u);
    raw += carry;
    uint32_t half = 1u << (g.width - 1);
    carry = raw > half ? 1u : 0u;
    return (int)raw - (int)(carry << g.width);
}

}  // namespace zk
