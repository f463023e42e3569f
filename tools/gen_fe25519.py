#!/usr/bin/env python3
"""Generator for the GF(2^255-19) multiply / square bodies used by the CUDA kernels.

Emits zkvm_b200/csrc/fe25519_mul.inc with two renderings of the same schedule:
  * device: inline-PTX carry chains (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses
    into IMAD.WIDE.U32[.X] with a predicate carry on sm_100a);
  * host:   the identical chain structure on unsigned __int128, compiled only by the
    CPU unit tests in tests/ (tests/host_emul) to validate the schedule without a GPU.

Schedule ("even/odd" wide-product columns): a 256x256 product is the sum of 64
32x32->64 partial products p(i,j)=a_j*b_i at limb position i+j.  Products whose
position has the same parity never overlap inside one row when j steps by 2, so each
row contributes two carry chains of four 64-bit slots.  Even-position chains go to
accumulator E (E[k] = limb k), odd-position chains to accumulator O (O[k] = limb k+1).
E + (O << 32) is the 512-bit product; it is folded with 2^256 = 38 (mod p).
"""
import os
import sys

# Chains over still-empty accumulator slots are emitted as independent wide multiplies (measured on B200: bucket
# accumulation -1.0 %, decoder -1.7 %); ZK_GEN_PLAIN_HEADS=0 regenerates the fully carry-chained form for A/B runs.
PLAIN_HEADS = os.environ.get("ZK_GEN_PLAIN_HEADS", "1") == "1"

class Acc:
    def __init__(self, name, n):
        self.name, self.n = name, n
        self.touched = [False] * n
    def ref(self, k): return f"{self.name}[{k}]"

class Emit:
    def __init__(self):
        self.dev, self.host = [], []

    def chain(self, acc, s, prods):
        """acc[s .. s+2*len(prods)-1] += sum prods[k] << (64*k); carry-out into acc[s+2*len]."""
        outs, ins = [], []          # asm operand lists
        def out_op(k, rw):          # returns %idx
            ref = acc.ref(k)
            for i, (r, m) in enumerate(outs):
                if r == ref: return i
            outs.append((ref, rw)); return len(outs) - 1
        lines, hl = [], []
        # decide operand modes first so numbering is stable: outputs first, then inputs
        slots = []
        for k, (x, y) in enumerate(prods):
            lo, hi = s + 2 * k, s + 2 * k + 1
            slots.append((lo, hi, x, y, acc.touched[lo], acc.touched[hi]))
        top = s + 2 * len(prods)
        need_carry = slots[-1][5] and top < acc.n   # top slot hi limb was live -> may overflow
        for (lo, hi, x, y, tl, th) in slots:
            out_op(lo, "+r" if tl else "=&r"); out_op(hi, "+r" if th else "=&r")
        if need_carry:
            out_op(top, "+r" if acc.touched[top] else "=&r")
        nout = len(outs)
        def in_op(expr):
            for i, e in enumerate(ins):
                if e == expr: return nout + i
            ins.append(expr); return nout + len(ins) - 1
        # A chain whose slots are all still empty adds nothing and carries nothing between its 64-bit products: emit
        # independent wide multiplies (IMAD.WIDE.U32 with a zero addend and no carry predicates issues faster than the
        # carry-chained form: tools/ubench/imad_rate.cu).
        if PLAIN_HEADS and not need_carry and all((not tl) and (not th) for (_, _, _, _, tl, th) in slots):
            dl = ["{ unsigned long long w_;"]
            hl.append("{ unsigned __int128 t_;")
            for (lo, hi, x, y, tl, th) in slots:
                dl.append(f'  asm("mul.wide.u32 %0, %1, %2;" : "=l"(w_) : "r"({x}), "r"({y})); {acc.ref(lo)} = (uint32_t)w_; {acc.ref(hi)} = (uint32_t)(w_ >> 32);')
                hl.append(f"  t_ = (unsigned __int128)({x}) * ({y}); {acc.ref(lo)} = (uint32_t)t_; {acc.ref(hi)} = (uint32_t)(t_ >> 32);")
                acc.touched[lo] = acc.touched[hi] = True
            dl.append("}"); hl.append("}")
            self.dev.append("\n".join(dl)); self.host += hl
            return
        first = True
        hl.append("{ unsigned __int128 t_; uint64_t c_ = 0;")
        for (lo, hi, x, y, tl, th) in slots:
            ol, oh = out_op(lo, None), out_op(hi, None)
            ix, iy = in_op(x), in_op(y)
            al = f"%{ol}" if tl else "0"
            ah = f"%{oh}" if th else "0"
            op_lo = "mad.lo.cc.u32" if first else "madc.lo.cc.u32"
            lines.append(f"{op_lo} %{ol}, %{ix}, %{iy}, {al};")
            lines.append(f"madc.hi.cc.u32 %{oh}, %{ix}, %{iy}, {ah};")
            first = False
            hl.append(f"  t_ = (unsigned __int128)({x}) * ({y}) + c_"
                      + (f" + {acc.ref(lo)}" if tl else "")
                      + (f" + ((uint64_t){acc.ref(hi)} << 32)" if th else "") + ";")
            hl.append(f"  {acc.ref(lo)} = (uint32_t)t_; {acc.ref(hi)} = (uint32_t)(t_ >> 32); c_ = (uint64_t)(t_ >> 64);")
            acc.touched[lo] = acc.touched[hi] = True
        if need_carry:
            ot = out_op(top, None)
            at = f"%{ot}" if acc.touched[top] else "0"
            lines.append(f"addc.u32 %{ot}, {at}, 0;")
            hl.append(f"  {acc.ref(top)} = " + (f"{acc.ref(top)} + " if acc.touched[top] else "") + "(uint32_t)c_;")
            acc.touched[top] = True
        hl.append("}")
        o = ", ".join(f'"{m}"({r})' for r, m in outs)
        i = ", ".join(f'"r"({e})' for e in ins)
        self.dev.append('asm("' + " ".join(lines) + f'" : {o} : {i});')
        self.host += hl

    def raw(self, dev, host):
        self.dev.append(dev); self.host.append(host)


def gen_fold(em):
    """r[0..15] (512-bit) -> out[0..7] = r mod-ish p (value < 2^256, congruent mod p)."""
    R = Acc("r", 17)
    R.touched = [True] * 16 + [False]
    # lo += 38*hi_even ; carry -> c1 ; then odd
    em.raw("uint32_t k38 = 38u, c1, l8;", "uint32_t k38 = 38u, c1, l8;")
    # even hi limbs at slots (0,1)..(6,7)
    dev = []
    dev.append('asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;"')
    dev.append('    "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;"')
    dev.append('    "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;"')
    dev.append('    "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.cc.u32 %7, %12, %13, %7;"')
    dev.append('    "addc.u32 %8, 0, 0;"')
    dev.append('    : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(c1)')
    dev.append('    : "r"(r[8]), "r"(r[10]), "r"(r[12]), "r"(r[14]), "r"(k38));')
    host = ["{ unsigned __int128 t_; uint64_t c_ = 0;"]
    for k in range(4):
        host.append(f"  t_ = (unsigned __int128)r[{8+2*k}] * 38u + c_ + r[{2*k}] + ((uint64_t)r[{2*k+1}] << 32);"
                    f" r[{2*k}] = (uint32_t)t_; r[{2*k+1}] = (uint32_t)(t_ >> 32); c_ = (uint64_t)(t_ >> 64);")
    host.append("  c1 = (uint32_t)c_; }")
    em.raw("\n".join(dev), "\n".join(host))
    # odd hi limbs at slots (1,2),(3,4),(5,6),(7,8): limb 8 starts as c1
    dev = []
    dev.append('asm("mad.lo.cc.u32 %0, %9, %13, %0; madc.hi.cc.u32 %1, %9, %13, %1;"')
    dev.append('    "madc.lo.cc.u32 %2, %10, %13, %2; madc.hi.cc.u32 %3, %10, %13, %3;"')
    dev.append('    "madc.lo.cc.u32 %4, %11, %13, %4; madc.hi.cc.u32 %5, %11, %13, %5;"')
    dev.append('    "madc.lo.cc.u32 %6, %12, %13, %6; madc.hi.u32 %7, %12, %13, %8;"')
    dev.append('    : "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(l8)')
    dev.append('    : "r"(c1), "r"(r[9]), "r"(r[11]), "r"(r[13]), "r"(r[15]), "r"(k38));')
    host = ["{ unsigned __int128 t_; uint64_t c_ = 0;"]
    for k in range(3):
        host.append(f"  t_ = (unsigned __int128)r[{9+2*k}] * 38u + c_ + r[{1+2*k}] + ((uint64_t)r[{2+2*k}] << 32);"
                    f" r[{1+2*k}] = (uint32_t)t_; r[{2+2*k}] = (uint32_t)(t_ >> 32); c_ = (uint64_t)(t_ >> 64);")
    host.append("  t_ = (unsigned __int128)r[15] * 38u + c_ + r[7] + ((uint64_t)c1 << 32);"
                " r[7] = (uint32_t)t_; l8 = (uint32_t)(t_ >> 32); }")
    em.raw("\n".join(dev), "\n".join(host))
    # fold limb 8 (<= ~40): r[0..7] += 38*l8, then the (rare) final carry once more
    dev = []
    dev.append('asm("mad.lo.cc.u32 %0, %9, %10, %0; addc.cc.u32 %1, %1, 0; addc.cc.u32 %2, %2, 0; addc.cc.u32 %3, %3, 0;"')
    dev.append('    "addc.cc.u32 %4, %4, 0; addc.cc.u32 %5, %5, 0; addc.cc.u32 %6, %6, 0; addc.cc.u32 %7, %7, 0; addc.u32 %8, 0, 0;"')
    dev.append('    : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "=&r"(c1)')
    dev.append('    : "r"(l8), "r"(k38));')
    dev.append("r[0] += c1 * 38u;")
    host = ["{ uint64_t c_ = (uint64_t)l8 * 38u;"]
    for k in range(8):
        host.append(f"  c_ += r[{k}]; r[{k}] = (uint32_t)c_; c_ >>= 32;")
    host.append("  r[0] += (uint32_t)c_ * 38u; }")
    em.raw("\n".join(dev), "\n".join(host))
    for k in range(8):
        em.raw(f"out[{k}] = r[{k}];", f"out[{k}] = r[{k}];")


def gen_merge(em, doubled=False):
    """r[0..15] = E + (O << 32)."""
    dev = ['asm("add.cc.u32 %0, %0, %15;"']
    for k in range(2, 15):
        dev.append(f'    "addc.cc.u32 %{k-1}, %{k-1}, %{14+k};"')
    dev.append('    "addc.u32 %14, %14, %29;"')
    dev.append("    : " + ", ".join(f'"+r"(E[{k}])' for k in range(1, 16)))
    dev.append("    : " + ", ".join(f'"r"(O[{k}])' for k in range(0, 15)) + ");")
    host = ["{ uint64_t c_ = 0;"]
    for k in range(1, 16):
        host.append(f"  c_ += (uint64_t)E[{k}] + O[{k-1}]; E[{k}] = (uint32_t)c_; c_ >>= 32;")
    host.append("}")
    em.raw("\n".join(dev), "\n".join(host))


def gen_mul():
    em = Emit()
    em.raw("uint32_t E[16], O[15];", "uint32_t E[16], O[15];")
    E, O = Acc("E", 16), Acc("O", 15)
    ev = lambda i: [(f"a[{j}]", f"b[{i}]") for j in (0, 2, 4, 6)]
    od = lambda i: [(f"a[{j}]", f"b[{i}]") for j in (1, 3, 5, 7)]
    for i in range(8):
        if i % 2 == 0:
            em.chain(E, i, ev(i)); em.chain(O, i, od(i))
        else:
            em.chain(O, i - 1, ev(i)); em.chain(E, i + 1, od(i))
    assert all(E.touched) and all(O.touched), (E.touched, O.touched)
    gen_merge(em)
    em.raw("uint32_t* r = E;", "uint32_t* r = E;")
    gen_fold(em)
    return em


def gen_sqr():
    em = Emit()
    em.raw("uint32_t E[16], O[15];", "uint32_t E[16], O[15];")
    E, O = Acc("E", 16), Acc("O", 15)
    # cross terms a_i*a_j, i<j, at position i+j
    for i in range(7):
        odd_js = [j for j in range(i + 1, 8, 2)]     # position 2i+1, 2i+3, ... (odd) -> O index pos-1
        even_js = [j for j in range(i + 2, 8, 2)]    # position 2i+2, ... (even) -> E
        if odd_js:
            em.chain(O, 2 * i, [(f"a[{i}]", f"a[{j}]") for j in odd_js])
        if even_js:
            em.chain(E, 2 * i + 2, [(f"a[{i}]", f"a[{j}]") for j in even_js])
    # untouched limbs are zero
    for k in range(16):
        if not E.touched[k]: em.raw(f"E[{k}] = 0;", f"E[{k}] = 0;")
    for k in range(15):
        if not O.touched[k]: em.raw(f"O[{k}] = 0;", f"O[{k}] = 0;")
    gen_merge(em)
    # double (C < 2^511 so no bit is lost), then add the squares a_i^2 at slots (2i, 2i+1)
    dev = ['asm("add.cc.u32 %0, %0, %0;"']
    for k in range(1, 15):
        dev.append(f'    "addc.cc.u32 %{k}, %{k}, %{k};"')
    dev.append('    "addc.u32 %15, %15, %15;"')
    dev.append("    : " + ", ".join(f'"+r"(E[{k}])' for k in range(16)) + ");")
    host = ["{ uint32_t c_ = 0, n_;"]
    for k in range(16):
        host.append(f"  n_ = E[{k}] >> 31; E[{k}] = (E[{k}] << 1) | c_; c_ = n_;")
    host.append("}")
    em.raw("\n".join(dev), "\n".join(host))
    E.touched = [True] * 16
    dev = []
    for i in range(8):
        op = "mad.lo.cc.u32" if i == 0 else "madc.lo.cc.u32"
        oph = "madc.hi.cc.u32" if i < 7 else "madc.hi.u32"
        dev.append(f'    "{op} %{2*i}, %{16+i}, %{16+i}, %{2*i}; {oph} %{2*i+1}, %{16+i}, %{16+i}, %{2*i+1};"')
    dev[0] = "asm(" + dev[0].lstrip()
    dev.append("    : " + ", ".join(f'"+r"(E[{k}])' for k in range(16)))
    dev.append("    : " + ", ".join(f'"r"(a[{k}])' for k in range(8)) + ");")
    host = ["{ unsigned __int128 t_; uint64_t c_ = 0;"]
    for i in range(8):
        host.append(f"  t_ = (unsigned __int128)a[{i}] * a[{i}] + c_ + E[{2*i}] + ((uint64_t)E[{2*i+1}] << 32);"
                    f" E[{2*i}] = (uint32_t)t_; E[{2*i+1}] = (uint32_t)(t_ >> 32); c_ = (uint64_t)(t_ >> 64);")
    host.append("}")
    em.raw("\n".join(dev), "\n".join(host))
    em.raw("uint32_t* r = E;", "uint32_t* r = E;")
    gen_fold(em)
    return em


def main(path):
    import io, os
    mul, sqr = gen_mul(), gen_sqr()
    f = io.StringIO()
    f.write("// GENERATED by tools/gen_fe25519.py -- do not edit.\n")
    f.write("// out = a*b (resp. a*a) mod p, all operands 8x32-bit little-endian limbs, value in [0, 2^256).\n")
    for name, em, args in (("fe_mul_limbs", mul, "uint32_t* out, const uint32_t* a, const uint32_t* b"),
                           ("fe_sqr_limbs", sqr, "uint32_t* out, const uint32_t* a")):
        f.write(f"ZK_HD ZK_INLINE void {name}({args}) {{\n#if defined(__CUDA_ARCH__)\n")
        f.write("\n".join("  " + l for l in "\n".join(em.dev).split("\n")))
        f.write("\n#else\n")
        f.write("\n".join("  " + l for l in "\n".join(em.host).split("\n")))
        f.write("\n#endif\n}\n\n")
    text = f.getvalue()
    # only touch the file when the content changes: its mtime drives the rebuild of every CUDA object
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as out:
            out.write(text)

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "zkvm_b200/csrc/fe25519_mul.inc")
