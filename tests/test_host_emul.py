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
