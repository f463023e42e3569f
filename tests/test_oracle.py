"""The oracles (big-integer Python + C restatement) against the golden vectors.  CPU only.

This is what pins the oracle: RFC 9496 Appendix A (re-verified transcription) and libsodium 1.0.20
known answers -- the reference itself has no fixtures."""
import numpy as np
import pytest

from oracle import ristretto255_ref as ref

H = bytes.fromhex


def test_py_generator_multiples(rfc_vectors):
    acc = ref.Point.identity()
    for h in rfc_vectors["generator_multiples"]:
        assert acc.encode().hex() == h
        assert ref.decode(H(h)).encode().hex() == h
        acc = acc + ref.BASEPOINT
    assert len(rfc_vectors["generator_multiples"]) == 16


def test_py_bad_encodings(rfc_vectors):
    n = 0
    for cat, lst in rfc_vectors["bad_encodings"].items():
        for h in lst:
            assert ref.decode(H(h)) is None, (cat, h)
            n += 1
    assert n == 29


def test_py_derivation(rfc_vectors):
    for v in rfc_vectors["derivation"]:
        assert ref.from_uniform_bytes(H(v["sha512"])).encode().hex() == v["element"]


def test_py_vs_sodium(sodium_vectors):
    for h, e in sodium_vectors["from_hash"]:
        assert ref.from_uniform_bytes(H(h)).encode().hex() == e
    for k, p, r in sodium_vectors["scalarmult"][:20]:
        assert (ref.decode(H(p)) * int.from_bytes(H(k), "little")).encode().hex() == r
    for b, ok in sodium_vectors["validity"]:
        assert (ref.decode(H(b)) is not None) == ok
    for case in sodium_vectors["msm"][:6]:
        got = ref.msm_naive([H(s) for s in case["scalars"]], [H(p) for p in case["points"]])
        assert got.hex() == case["result"]


def test_c_generator_multiples_and_bad(rfc_vectors, c_oracle):
    B = H(rfc_vectors["generator_multiples"][1])
    for i, h in enumerate(rfc_vectors["generator_multiples"]):
        assert c_oracle.scalarmult(i.to_bytes(32, "little"), B).hex() == h
        assert c_oracle.is_valid(H(h))
    for cat, lst in rfc_vectors["bad_encodings"].items():
        for h in lst:
            assert not c_oracle.is_valid(H(h)), (cat, h)
            assert c_oracle.msm(bytes(32), H(h), 1) is None


def test_c_vs_sodium(sodium_vectors, c_oracle):
    blob = b"".join(H(h) for h, _ in sodium_vectors["from_hash"])
    got = c_oracle.from_uniform(blob, len(sodium_vectors["from_hash"]))
    assert got == b"".join(H(e) for _, e in sodium_vectors["from_hash"])
    for k, p, r in sodium_vectors["scalarmult"]:
        assert c_oracle.scalarmult(H(k), H(p)).hex() == r
    for b, ok in sodium_vectors["validity"]:
        assert c_oracle.is_valid(H(b)) == ok
    for case in sodium_vectors["msm"]:
        n = len(case["scalars"])
        s = b"".join(H(x) for x in case["scalars"]); p = b"".join(H(x) for x in case["points"])
        for threads in (1, 3):
            assert c_oracle.msm(s, p, n, threads).hex() == case["result"], n


@pytest.mark.parametrize("n,threads", [(150, 1), (190, 1), (700, 2), (1500, 1), (3000, 4)])
def test_c_msm_algorithms_agree_with_scalarmult_sum(c_oracle, n, threads):
    """Straus (n<190) and Pippenger (each digit width) against sum of independent double-and-add products."""
    rng = np.random.default_rng(n)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    prods = b"".join(c_oracle.scalarmult(bytes(sc[i]), pts[32 * i:32 * i + 32]) for i in range(n))
    assert c_oracle.msm(sc, pts, n, threads) == c_oracle.point_sum(prods, n)


def test_c_decompress_reports_first_bad_index(c_oracle, rfc_vectors):
    good = H(rfc_vectors["generator_multiples"][3])
    bad = H(rfc_vectors["bad_encodings"]["non_square"][0])
    blob = good * 10 + bad + good * 5 + bad
    _, idx = c_oracle.decompress(blob, 17, threads=1)
    assert idx == 10
    _, idx = c_oracle.decompress(good * 17, 17, threads=2)
    assert idx is None


def test_even_subgroup_criterion():
    """What the CUDA ingestion of extended points relies on (k_from_extended): a curve point lies in the even subgroup
    2E -- the set ristretto255 representatives are drawn from (RFC 9496 section 3) -- iff Z^2 - Y^2 is a square."""
    import random
    from oracle import ristretto255_ref as ref
    P, L = ref.P, ref.L
    rnd = random.Random(3)

    def smul(k, p):
        r = ref.Point.identity()
        while k:
            if k & 1: r = r + p
            p = p + p; k >>= 1
        return r

    def on_curve_point():
        while True:
            y = rnd.randrange(P)
            x2 = (y * y - 1) * pow(ref.D * y * y + 1, P - 2, P) % P
            x = pow(x2, (P + 3) // 8, P)
            if x * x % P != x2: x = x * ref.SQRT_M1 % P
            if x * x % P == x2: return ref.Point(x, y, 1, x * y)

    def is_square(a):
        a %= P
        return a == 0 or pow(a, (P - 1) // 2, P) == 1

    t8 = None
    while t8 is None:
        q = smul(L, on_curve_point())
        q4 = smul(4, q)
        if q4.X % P != 0: continue                      # (0, -1) has X = 0: q4 must be the order-2 point, q of order 8
        if (q4.Y - q4.Z) % P != 0: t8 = q
    for _ in range(25):
        base = on_curve_point().double()
        lam = rnd.randrange(1, P)
        for k in range(8):
            pt = base + smul(k, t8)
            pt = ref.Point(pt.X * lam, pt.Y * lam, pt.Z * lam, pt.T * lam)
            assert is_square(pt.Z * pt.Z - pt.Y * pt.Y) == (k % 2 == 0)
            if k % 2 == 0:
                assert pt.encode() == base.encode()     # adding 4-torsion stays in the same ristretto coset


def test_c_vector_backend_matches_serial(c_oracle, sodium_vectors):
    """The 4-lane AVX-512 IFMA bucket accumulation of the C port (timing arm only; dalek's vector-backend shape) gives the
    same bytes as the serial radix-2^51 code it stands beside: random and adversarial scalars, every Pippenger width,
    one and several threads, and the libsodium known answers.  Without IFMA the switch must simply refuse."""
    have = c_oracle.set_vector(True)
    c_oracle.set_vector(False)
    if not have:
        assert c_oracle.set_vector(True) is False
        pytest.skip("host CPU has no AVX-512 IFMA: the vector backend is not selectable here")
    L = 2**252 + 27742317777372353535851937790883648493
    rng = np.random.default_rng(31)

    def both(sc, pts, n, threads):
        c_oracle.set_vector(False); a = c_oracle.msm(sc, pts, n, threads=threads)
        try:
            assert c_oracle.set_vector(True); b = c_oracle.msm(sc, pts, n, threads=threads)
        finally:
            c_oracle.set_vector(False)
        assert a is not None and a == b, (n, threads)
        return a

    for n, threads in ((190, 1), (300, 1), (499, 2), (650, 1), (800, 1), (3000, 3), (40000, 4)):
        pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
        both(rng.integers(0, 256, size=(n, 32), dtype=np.uint8), pts, n, threads)
        # adversarial scalar sets: one bucket per column, all digits negative, zero, small, around l, unreduced maxima
        for v in (1, 2, L - 1, L - 2, (L - 1) // 2, int("80" * 32, 16) % L, int("7f" * 31 + "0f", 16), 0, 2**256 - 1, 8 * L + 3):
            sc = np.frombuffer((v % 2**256).to_bytes(32, "little") * n, dtype=np.uint8)
            both(sc, pts, n, threads)
        # repeated points (P + P inside a bucket: the unified addition must double correctly) and P, -P pairs
        rep = np.frombuffer(bytes(pts[:32]) * n, dtype=np.uint8)
        both(rng.integers(0, 256, size=(n, 32), dtype=np.uint8), rep, n, threads)
        sc = rng.integers(0, 256, size=(n // 2, 32), dtype=np.uint8)
        neg = np.array([(L - int.from_bytes(bytes(r), "little") % L) % L for r in sc], dtype=object)
        sc2 = np.concatenate([sc.reshape(-1), np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in neg), dtype=np.uint8)])
        half = np.frombuffer(bytes(pts[: 32 * (n // 2)]), dtype=np.uint8)
        r = both(sc2, np.concatenate([half, half]), 2 * (n // 2), threads)
        assert r == bytes(32)                                      # sum s_i P_i + sum (-s_i) P_i = identity
    for case in sodium_vectors["msm"]:
        n = len(case["scalars"])
        s = np.frombuffer(b"".join(H(x) for x in case["scalars"]), dtype=np.uint8)
        p = np.frombuffer(b"".join(H(x) for x in case["points"]), dtype=np.uint8)
        assert both(s, p, n, 1).hex() == case["result"]
