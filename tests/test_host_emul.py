"""The device headers (zkvm_b200/csrc/*.cuh) compiled for the host -- PTX carry chains replaced by their
portable rendering from the same generator -- checked against the big-integer oracle.  This validates the
limb schedules, reduction identities and the RFC 9496 routines the kernels run, without a GPU.  CPU only."""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle import ristretto255_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ref.P


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libemul.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests/host_emul/emul.cpp")])
    return C.CDLL(so)


def test_generated_file_is_current(tmp_path):
    out = tmp_path / "fe.inc"
    subprocess.check_call(["python", os.path.join(ROOT, "tools/gen_fe25519.py"), str(out)])
    assert out.read_text() == open(os.path.join(ROOT, "zkvm_b200/csrc/fe25519_mul.inc")).read()


def _run(emul, op, A, B):
    n = len(A)
    ab = b"".join(x.to_bytes(32, "little") for x in A); bb = b"".join(x.to_bytes(32, "little") for x in B)
    out = C.create_string_buffer(32 * n)
    emul.emul_fe_op(op, ab, bb, out, C.c_size_t(n))
    return [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(n)]


EDGE = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**256 - 1, 2**255, 2**255 - 1, 2**128, 2**256 - 38,
        2**256 - 39, 0xffffffff, 2**224 - 1, 2**32, 2**64 - 1, (2**256 - 1) // 3]


def test_field_ops(emul):
    rnd = random.Random(1)
    vals = EDGE + [rnd.getrandbits(256) for _ in range(2000)]
    A = [rnd.choice(vals) for _ in range(4000)] + [a for a in EDGE for _ in EDGE]
    B = [rnd.choice(vals) for _ in range(4000)] + [b for _ in EDGE for b in EDGE]
    for op, f in ((0, lambda a, b: a * b), (1, lambda a, b: a * a), (2, lambda a, b: a + b), (3, lambda a, b: a - b),
                  (6, lambda a, b: -a)):
        for a, b, x in zip(A, B, _run(emul, op, A, B)):
            assert (x - f(a, b)) % P == 0, (op, hex(a), hex(b))
            assert 0 <= x < 2**256
    for a, x in zip(A, _run(emul, 4, A, B)):
        assert x == a % P
    for a, x in zip(A[:100], _run(emul, 5, A[:100], B[:100])):
        assert (x - pow(a, P - 2, P)) % P == 0
    for a, x in zip(A[:100], _run(emul, 7, A[:100], B[:100])):
        assert (x - pow(a, (P - 5) // 8, P)) % P == 0


def test_ristretto_codec(emul, rfc_vectors, sodium_vectors):
    xyt = C.create_string_buffer(96); o = C.create_string_buffer(32)
    for h in rfc_vectors["generator_multiples"]:
        b = bytes.fromhex(h)
        assert emul.emul_decode(b, xyt) == 1
        pt = ref.decode(b)
        x = int.from_bytes(xyt.raw[:32], "little") % P; y = int.from_bytes(xyt.raw[32:64], "little") % P
        assert (x, y) == (pt.X, pt.Y)
        ext = b"".join(v.to_bytes(32, "little") for v in (x, y, 1, x * y % P))
        emul.emul_encode(ext, o)
        assert o.raw == b
    for lst in rfc_vectors["bad_encodings"].values():
        for h in lst:
            assert emul.emul_decode(bytes.fromhex(h), xyt) == 0, h
    for b, ok in sodium_vectors["validity"]:
        assert bool(emul.emul_decode(bytes.fromhex(b), xyt)) == ok
    for h, e in sodium_vectors["from_hash"]:
        emul.emul_from_uniform(bytes.fromhex(h), o)
        assert o.raw.hex() == e
    # encode must be representative-independent: scale (X:Y:Z:T) by a random factor and add 4-torsion
    rnd = random.Random(5)
    for h in rfc_vectors["generator_multiples"][1:6]:
        pt = ref.decode(bytes.fromhex(h))
        lam = rnd.getrandbits(250) + 2
        for q in (pt, ref.Point(pt.Y * ref.SQRT_M1, pt.X * ref.SQRT_M1, pt.Z, -pt.T), ref.Point(-pt.X, -pt.Y, pt.Z, pt.T)):
            ext = b"".join((v * lam % P).to_bytes(32, "little") for v in (q.X, q.Y, q.Z, q.T))
            emul.emul_encode(ext, o)
            assert o.raw.hex() == h


def test_group_law_msm(emul, sodium_vectors):
    for case in sodium_vectors["msm"][:5]:
        n = len(case["scalars"])
        s = b"".join(bytes.fromhex(x) for x in case["scalars"]); p = b"".join(bytes.fromhex(x) for x in case["points"])
        o = C.create_string_buffer(32)
        assert emul.emul_msm(s, p, C.c_size_t(n), o) == 0
        assert o.raw.hex() == case["result"]


def test_scalar_recoding(emul):
    """recode.cuh (the code the sort kernels run): reduction mod l of any 256-bit string, balanced window geometry
    covering 254 bits, signed digits within half a window, and  sum_w d_w 2^(off_w) == s mod l  with no carry left --
    for every window count the library can pick or be forced to (c = 4..20) and edge scalars around l and 2^256."""
    L = 2**252 + 27742317777372353535851937790883648493
    rnd = random.Random(9)
    edge = [0, 1, 2, L - 1, L, L + 1, 2 * L, 8 * L, 8 * L - 1, 15 * L + 5, 2**252 - 1, 2**252, 2**253 - 1, 2**253, 2**255, 2**256 - 1,
            (2**256 - 1) // 3, 2**128, 2**128 - 1] + [(1 << k) - 1 for k in range(1, 256, 17)] + [1 << k for k in range(0, 256, 13)]
    # all-ones windows provoke the longest carry runs; 0x8000.. patterns sit exactly on the half-window boundary
    edge += [int("8" + "0" * 63, 16) % 2**256, int("7f" * 32, 16), int("80" * 32, 16), int("ff" * 32, 16) % L, L // 2, L // 2 + 1]
    scalars = edge + [rnd.getrandbits(256) for _ in range(300)] + [rnd.getrandbits(253) for _ in range(100)]
    red = C.create_string_buffer(32)
    for c in range(4, 21):
        W = (254 + c - 1) // c
        dig = (C.c_int32 * (W + 1))(); off = (C.c_int32 * W)(); wid = (C.c_int32 * W)()
        for s in scalars:
            emul.emul_recode(s.to_bytes(32, "little"), W, red, dig, off, wid)
            assert int.from_bytes(red.raw, "little") == s % L
            assert dig[W] == 0, (c, hex(s))
            assert off[0] == 0 and all(off[w + 1] == off[w] + wid[w] for w in range(W - 1)) and off[W - 1] + wid[W - 1] == 254
            assert max(wid) - min(wid) <= 1 and max(wid) == (254 + W - 1) // W
            assert all(abs(dig[w]) <= 1 << (wid[w] - 1) for w in range(W)), (c, hex(s))
            assert sum(dig[w] << off[w] for w in range(W)) == s % L, (c, hex(s))
