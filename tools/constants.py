#!/usr/bin/env python3
"""Derive the curve constants embedded in zkvm_b200/csrc/fe25519.cuh (RFC 9496 section 4.1)."""
p = 2**255 - 19
d = (-121665 * pow(121666, p - 2, p)) % p
sqrt_m1 = pow(2, (p - 1) // 4, p)

def limbs(x):
    return ", ".join(f"0x{(x >> (32 * i)) & 0xffffffff:08x}u" for i in range(8))

def sqrt_ratio(u, v):
    v3 = v * v % p * v % p; v7 = v3 * v3 % p * v % p
    r = u * v3 % p * pow(u * v7 % p, (p - 5) // 8, p) % p
    check = v * r % p * r % p
    cs = check == u % p; fl = check == (-u) % p; fli = check == (-u * sqrt_m1) % p
    if fl or fli: r = r * sqrt_m1 % p
    if r & 1: r = p - r
    return (cs or fl), r

ok, invsqrt_a_minus_d = sqrt_ratio(1, (-1 - d) % p); assert ok
ok, s = sqrt_ratio((-d - 1) % p, 1); assert ok
sqrt_ad_minus_one = p - s          # RFC 9496 lists the odd root
RFC = {
    "D": 37095705934669439343138083508754565189542113879843219016388785533085940283555,
    "SQRT_M1": 19681161376707505956807079304988542015446066515923890162744021073123829784752,
    "SQRT_AD_MINUS_ONE": 25063068953384623474111414158702152701244531502492656460079210482610430750235,
    "INVSQRT_A_MINUS_D": 54469307008909316920995813868745141605393597292927456921205312896311721017578,
    "ONE_MINUS_D_SQ": 1159843021668779879193775521855586647937357759715417654439879720876111806838,
    "D_MINUS_ONE_SQ": 40440834346308536858101042469323190826248399146238708352240133220865137265952,
}
vals = {"D": d, "D2": 2 * d % p, "SQRT_M1": sqrt_m1, "ONE_MINUS_D_SQ": (1 - d * d) % p,
        "D_MINUS_ONE_SQ": (d - 1) ** 2 % p, "INVSQRT_A_MINUS_D": invsqrt_a_minus_d,
        "SQRT_AD_MINUS_ONE": sqrt_ad_minus_one}
for k, v in vals.items():
    tag = ""
    if k in RFC:
        tag = "  [matches RFC 9496 4.1]" if RFC[k] == v else "  [MISMATCH vs recalled RFC value]"
    print(f"{k:20s} {{{limbs(v)}}}{tag}")
