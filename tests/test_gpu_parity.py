"""Parity of the CUDA path (through the C ABI) against the oracle, the golden vectors, and -- at the full
BASELINE.json sizes -- size-independent properties.  Bit-exact: every comparison is on bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
H = bytes.fromhex
L = 2**252 + 27742317777372353535851937790883648493


def le32(v):
    return (v % 2**256).to_bytes(32, "little")


def make_points(c_oracle, n, seed):
    rng = np.random.default_rng(seed)
    return c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)


def rand_scalars(n, seed):
    return np.random.default_rng(seed ^ 0xabc).integers(0, 256, size=(n, 32), dtype=np.uint8)


# ---------------- codec ----------------
def test_decode_encode_golden(ctx, rfc_vectors):
    import zkvm_b200 as zk
    enc = b"".join(H(h) for h in rfc_vectors["generator_multiples"])
    t = zk.PointTable(ctx).append_compressed(enc)
    assert len(t) == 16
    assert t.compress() == enc
    assert t.compress(offset=3, n=2) == enc[96:160]


def test_bad_encodings_rejected_with_index(ctx, rfc_vectors):
    import zkvm_b200 as zk
    good = H(rfc_vectors["generator_multiples"][5])
    for cat, lst in rfc_vectors["bad_encodings"].items():
        for k, h in enumerate(lst):
            t = zk.PointTable(ctx)
            blob = good * (k + 2) + H(h) + good * 3 + H(h)
            with pytest.raises(zk.InvalidPoint) as e:
                t.append_compressed(blob)
            assert e.value.index == k + 2, (cat, h)
            assert len(t) == 0
            assert zk.CompressedRistretto(H(h)).decompress(ctx) is None
            assert zk.RistrettoPoint.optional_multiscalar_mul(ctx, le32(1) * 2, good + H(h)) is None


def test_validity_matches_sodium(ctx, sodium_vectors):
    import zkvm_b200 as zk
    for b, ok in sodium_vectors["validity"]:
        assert (zk.CompressedRistretto(H(b)).decompress(ctx) is not None) == ok


def test_from_uniform_golden_and_oracle(ctx, rfc_vectors, sodium_vectors, c_oracle):
    import zkvm_b200 as zk
    t = zk.PointTable(ctx).append_uniform(b"".join(H(h) for h, _ in sodium_vectors["from_hash"]))
    assert t.compress() == b"".join(H(e) for _, e in sodium_vectors["from_hash"])
    for v in rfc_vectors["derivation"]:
        assert zk.PointTable(ctx).append_uniform(H(v["sha512"])).compress().hex() == v["element"]
    u = np.random.default_rng(11).integers(0, 256, size=(5000, 64), dtype=np.uint8)
    t = zk.PointTable(ctx, 5000).append_uniform(u)
    enc = t.compress()
    assert enc == c_oracle.from_uniform(u, 5000)
    assert zk.PointTable(ctx).append_compressed(enc).compress() == enc      # decode(encode(P)) == P


# ---------------- MSM vs golden / oracle ----------------
def test_msm_sodium_golden(ctx, sodium_vectors):
    import zkvm_b200 as zk
    for case in sodium_vectors["msm"]:
        s = b"".join(H(x) for x in case["scalars"]); p = b"".join(H(x) for x in case["points"])
        for c in (0, 4, 9, 16):
            ctx.set_window(c)
            got = zk.RistrettoPoint.optional_multiscalar_mul(ctx, s, p)
            assert bytes(got).hex() == case["result"], (len(case["scalars"]), c)
    ctx.set_window(0)


def test_scalarmult_golden_edge_scalars(ctx, sodium_vectors):
    """n = 1 MSMs with scalars 0, 1, l-1, l, l+1, 2^255, 2^256-1, ... (unreduced scalars are taken mod l)."""
    import zkvm_b200 as zk
    for k, p, r in sodium_vectors["scalarmult"]:
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, H(k), H(p))).hex() == r


@pytest.mark.parametrize("n", [1, 2, 31, 189, 190, 1000, 4097, 1 << 14])
def test_msm_vs_oracle_all_windows(ctx, c_oracle, n):
    import zkvm_b200 as zk
    pts = make_points(c_oracle, n, n); sc = rand_scalars(n, n)
    want = c_oracle.msm(sc, pts, n, threads=4)
    tab = zk.PointTable(ctx, n).append_compressed(pts)
    for c in (0, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16):
        ctx.set_window(c)
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want, c
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want, c
    ctx.set_window(0)


def test_msm_adversarial_inputs(ctx, c_oracle):
    """All-identity points, one repeated point, scalars 0 / 1 / l-1, everything in one bucket."""
    import zkvm_b200 as zk
    n = 3000
    pts = make_points(c_oracle, n, 99)
    cases = {
        "all identity points": (rand_scalars(n, 1).tobytes(), bytes(32) * n),
        "repeated point": (rand_scalars(n, 2).tobytes(), pts[:32] * n),
        "scalars all zero": (bytes(32 * n), pts),
        "scalars all one": (le32(1) * n, pts),
        "scalars all l-1": (le32(L - 1) * n, pts),
        "scalars all equal random": (le32(0x1234567890abcdef << 100 | 77) * n, pts),
        "scalars 0,1,l-1 cycling": (b"".join(le32(v) for v in (0, 1, L - 1)) * (n // 3), pts),
        "single high window": (le32(1 << 251) * n, pts),
    }
    for name, (s, p) in cases.items():
        want = c_oracle.msm(s, p, n, threads=4)
        for c in (0, 8, 16):
            ctx.set_window(c)
            assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, s, p)) == want, (name, c)
    ctx.set_window(0)
    assert want is not None
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, b"", b"")) == bytes(32)     # empty MSM = identity


def test_skewed_digit_distributions(ctx, c_oracle):
    """Scalars that pile entries into few buckets (one window's digit identical for every term, half the terms equal,
    all equal) at sizes where a bucket is split into many tasks: the counting sort and the task planner are exact for any
    distribution."""
    import zkvm_b200 as zk
    for n in (33, 3000, 70000):
        pts = make_points(c_oracle, n, 600 + n)
        tab = zk.PointTable(ctx).append_compressed(pts)
        skew = rand_scalars(n, 601 + n).copy(); skew[:, 7] = 0x5a; skew[:, 8] = 0xc3          # bits 56..71 equal for all
        half = rand_scalars(n, 602 + n).copy(); half[: n // 2] = half[0]                       # half the terms identical
        for name, sc in {"one_hot_window": skew, "half_equal": half, "all_equal": np.tile(rand_scalars(1, 9), (n, 1))}.items():
            want = c_oracle.msm(sc, pts, n, threads=4)
            for c in (0, 6, 11):
                ctx.set_window(c)
                try:
                    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want, (n, name, c)
                finally:
                    ctx.set_window(0)
            assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want, (n, name)


def test_sum_compressed(ctx, c_oracle, rfc_vectors):
    """zk_sum_compressed: the combine step of a one-process-per-GPU deployment, against the oracle's point sum."""
    import zkvm_b200 as zk
    pts = make_points(c_oracle, 1024, 4242)
    for g in (0, 1, 2, 7, 255, 256, 257, 300, 1024):
        got = ctx.sum_compressed(pts[:32 * g])
        want = c_oracle.point_sum(pts[:32 * g], g) if g else bytes(32)
        assert got is not None and bytes(got) == want, g
    bad = bytearray(pts[:32 * 300]); bad[32 * 299:32 * 300] = H(rfc_vectors["bad_encodings"]["negative_s"][0])
    assert ctx.sum_compressed(bytes(bad)) is None
    with pytest.raises(zk.ZkError):
        ctx.sum_compressed(pts + pts[:32])            # more than 1024 points
    hp = zk.Context(0); hp.set_priority(True)         # a high-priority context computes the same bytes
    assert bytes(hp.sum_compressed(pts[:32 * 9])) == c_oracle.point_sum(pts[:32 * 9], 9)
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(hp, rand_scalars(100, 1), pts[:3200])) == c_oracle.msm(rand_scalars(100, 1), pts[:3200], 100)
    hp.close()


def test_invalid_point_anywhere_rejects(ctx, c_oracle, rfc_vectors):
    import zkvm_b200 as zk
    n = 2048
    pts = bytearray(make_points(c_oracle, n, 5)); sc = rand_scalars(n, 5)
    bad = H(rfc_vectors["bad_encodings"]["negative_xy"][2])
    for idx in (0, 777, n - 1):
        q = bytearray(pts); q[32 * idx:32 * idx + 32] = bad
        assert zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, bytes(q)) is None
        assert c_oracle.msm(sc, bytes(q), n) is None
    assert zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, bytes(pts)) is not None


def test_table_slices_and_mixed(ctx, c_oracle):
    """Cached static prefix + dynamic compressed suffix (the bulletproofs verification shape)."""
    import zkvm_b200 as zk
    n_s, n_d = 1500, 300
    gens = make_points(c_oracle, 4000, 21); dyn = make_points(c_oracle, n_d, 22)
    s_s = rand_scalars(n_s, 23); s_d = rand_scalars(n_d, 24)
    tab = zk.PointTable(ctx).append_compressed(gens[:32 * 1000]).append_compressed(gens[32 * 1000:])   # growth path
    assert len(tab) == 4000
    off = 700
    want = c_oracle.msm(s_s.tobytes() + s_d.tobytes(), gens[32 * off:32 * (off + n_s)] + dyn, n_s + n_d, threads=4)
    got = zk.RistrettoPoint.mixed_multiscalar_mul(ctx, s_s, tab, s_d, dyn, offset=off)
    assert bytes(got) == want
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, s_s, tab, offset=off)) == \
        c_oracle.msm(s_s, gens[32 * off:32 * (off + n_s)], n_s, threads=4)
    # only static / only dynamic
    assert bytes(zk.RistrettoPoint.mixed_multiscalar_mul(ctx, s_s, tab, b"", b"", offset=off)) == \
        c_oracle.msm(s_s, gens[32 * off:32 * (off + n_s)], n_s, threads=4)
    assert bytes(zk.RistrettoPoint.mixed_multiscalar_mul(ctx, b"", None, s_d, dyn)) == c_oracle.msm(s_d, dyn, n_d)
    with pytest.raises(zk.ZkError):
        zk.RistrettoPoint.vartime_multiscalar_mul(ctx, s_s, tab, offset=3000)       # slice out of range


# ---------------- full-size properties (n = 2^20, the headline config) ----------------
@pytest.fixture(scope="module")
def big(ctx):
    import zkvm_b200 as zk
    n = 1 << 20
    u = np.random.default_rng(2020).integers(0, 256, size=(n, 64), dtype=np.uint8)
    tab = zk.PointTable(ctx, n).append_uniform(u)
    return n, tab, tab.compress(), rand_scalars(n, 2020)


def test_big_window_independence_and_compressed_path(ctx, big):
    import zkvm_b200 as zk
    n, tab, comp, sc = big
    res = set()
    for c in (13, 15, 16, 0):
        ctx.set_window(c)
        res.add(bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)))
    ctx.set_window(0)
    res.add(bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, comp)))     # decompress-on-device path
    assert len(res) == 1


def test_big_negation_cancels(ctx, big):
    """sum s_i P_i + sum (l - s_i) P_i = identity: one 2^21-point MSM that must encode to 32 zero bytes."""
    import zkvm_b200 as zk
    n, tab, comp, sc = big
    half = 1 << 19
    s_int = [int.from_bytes(bytes(r), "little") % L for r in sc[:half]]
    neg = b"".join(le32(L - v) for v in s_int)
    out = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc[:half].tobytes() + neg, comp[:32 * half] * 2)
    assert out.is_identity()


def test_big_linearity_against_oracle_scalarmult(ctx, big, c_oracle):
    """MSM(k * 1, P) == k * MSM(1, P): the right side is one oracle scalar multiplication of the GPU's point sum."""
    import zkvm_b200 as zk
    n, tab, comp, sc = big
    total = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, le32(1) * n, tab)
    k = 0x0123456789abcdef0123456789abcdef0123456789abcdef0123456789abcdef % L
    got = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, le32(k) * n, tab)
    assert bytes(got) == c_oracle.scalarmult(le32(k), bytes(total))


def test_big_split_sums_to_whole(ctx, big, c_oracle):
    """Index-range partial MSMs add up to the whole (the multi-GPU sharding identity), checked by the oracle."""
    import zkvm_b200 as zk
    n, tab, comp, sc = big
    whole = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
    cuts = [0, 100000, 1 << 19, n - 3, n]
    parts = b"".join(bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc[a:b], tab, offset=a)) for a, b in zip(cuts, cuts[1:]))
    assert c_oracle.point_sum(parts, len(cuts) - 1) == bytes(whole)


def test_sample_of_big_against_oracle(ctx, big, c_oracle):
    import zkvm_b200 as zk
    n, tab, comp, sc = big
    m = 1 << 15
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc[5000:5000 + m], tab, offset=5000)) == \
        c_oracle.msm(sc[5000:5000 + m], comp[32 * 5000:32 * (5000 + m)], m, threads=8)


def test_full_2e20_against_oracle(ctx, big, c_oracle):
    """The headline size compared DIRECTLY with the CPU oracle on identical bytes: table path, compressed path and
    window-expanded path (the multi-threaded C oracle needs well under a second per 2^20 terms)."""
    import os
    import zkvm_b200 as zk
    n, tab, comp, sc = big
    want = c_oracle.msm(sc, comp, n, threads=min(32, os.cpu_count() or 1))
    assert want is not None
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, comp)) == want
    ctx.set_staging(2)                      # the same call with every upload forced through the pinned staging ring
    try:
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, comp)) == want
    finally:
        ctx.set_staging(0)
    tab.precompute(0)
    try:
        assert tab.precomputed_window > 0
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want
    finally:
        tab.clear(); tab.append_compressed(comp)


# ---------------- batches of independent MSMs ----------------
def test_batch_matches_individual_msms(ctx, c_oracle, rfc_vectors):
    import zkvm_b200 as zk
    sizes = [0, 1, 5, 300, 0, 1000, 189, 2048, 64, 7]
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    n = int(seg[-1])
    pts = make_points(c_oracle, n, 314); sc = rand_scalars(n, 314)
    want = [c_oracle.msm(sc[int(a):int(b)], pts[32 * int(a):32 * int(b)], int(b - a), threads=2) for a, b in zip(seg, seg[1:])]
    for c in (0, 5, 8, 13):
        ctx.set_window(c)
        got = zk.batch_optional_multiscalar_mul(ctx, sc, pts, seg)
        assert [bytes(g) for g in got] == want, c
    ctx.set_window(0)
    assert want[0] == bytes(32) and want[4] == bytes(32)                    # empty MSM = identity
    # one invalid encoding voids only its own MSM
    bad = bytearray(pts); k = int(seg[5]) + 17
    bad[32 * k:32 * k + 32] = H(rfc_vectors["bad_encodings"]["non_square"][1])
    got = zk.batch_optional_multiscalar_mul(ctx, sc, bytes(bad), seg)
    assert got[5] is None
    assert [bytes(g) for i, g in enumerate(got) if i != 5] == [w for i, w in enumerate(want) if i != 5]


def test_batch_many_small(ctx, c_oracle):
    """1024 'transactions' of 256 terms each, every one with its own result."""
    import zkvm_b200 as zk
    m, per = 1024, 256
    n = m * per
    pts = make_points(c_oracle, n, 2718); sc = rand_scalars(n, 2718)
    seg = np.arange(0, n + 1, per, dtype=np.uint64)
    got = zk.batch_optional_multiscalar_mul(ctx, sc, pts, seg)
    for k in (0, 1, 511, 1023):
        assert bytes(got[k]) == c_oracle.msm(sc[k * per:(k + 1) * per], pts[32 * k * per:32 * (k + 1) * per], per)
    # the batch adds up to the one big MSM over everything
    assert c_oracle.point_sum(b"".join(bytes(g) for g in got), m) == bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts))


def test_batch_over_shared_table(ctx, c_oracle):
    """m proofs against one set of cached generators: MSM k uses table[offset : offset + len_k]."""
    import zkvm_b200 as zk
    gens = make_points(c_oracle, 3000, 1618)
    tab = zk.PointTable(ctx).append_compressed(gens)
    sizes = [2500, 0, 1, 777, 2500, 64]
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    sc = rand_scalars(int(seg[-1]), 1618)
    off = 123
    got = zk.batch_vartime_multiscalar_mul(ctx, sc, tab, seg, offset=off)
    for k, (a, b) in enumerate(zip(seg, seg[1:])):
        a, b = int(a), int(b)
        assert bytes(got[k]) == c_oracle.msm(sc[a:b], gens[32 * off:32 * (off + b - a)], b - a, threads=2), k
    with pytest.raises(zk.ZkError):
        zk.batch_vartime_multiscalar_mul(ctx, sc, tab, seg, offset=600)     # longest MSM would run past the table


def test_extended_point_ingestion(ctx, rfc_vectors):
    """Decompressed points handed over as (X, Y, Z, T): any projective representative, any coset member."""
    import random
    import zkvm_b200 as zk
    from oracle import ristretto255_ref as ref
    rnd = random.Random(9)
    P = ref.P
    blob, want = b"", b""
    for h in rfc_vectors["generator_multiples"]:
        pt = ref.decode(H(h))
        lam = rnd.getrandbits(250) + 2
        # a different member of the same ristretto coset (add 4-torsion), scaled projectively
        q = ref.Point(pt.Y * ref.SQRT_M1, pt.X * ref.SQRT_M1, pt.Z, -pt.T) if rnd.random() < 0.5 else pt
        blob += b"".join((v * lam % P).to_bytes(32, "little") for v in (q.X, q.Y, q.Z, q.T))
        want += H(h)
    t = zk.PointTable(ctx).append_extended(blob)
    assert t.compress() == want
    sc = rand_scalars(16, 4)
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, t)) == ref.msm_naive([bytes(r) for r in sc], [want[32 * i:32 * i + 32] for i in range(16)])
    good = blob[128:256]
    B = ref.BASEPOINT
    for name, bad in {
        "off curve": b"".join(v.to_bytes(32, "little") for v in (B.X, (B.Y + 1) % P, 1, B.X * (B.Y + 1) % P)),
        "T inconsistent": b"".join(v.to_bytes(32, "little") for v in (B.X, B.Y, 1, (B.T + 1) % P)),
        "Z zero": b"".join(v.to_bytes(32, "little") for v in (B.X, B.Y, 0, B.T)),
        "non-canonical coordinate": b"".join(v.to_bytes(32, "little") for v in (B.X + P, B.Y, 1, B.T)),
    }.items():
        tt = zk.PointTable(ctx)
        with pytest.raises(zk.InvalidPoint) as e:
            tt.append_extended(good * 3 + bad + good)
        assert e.value.index == 3 and len(tt) == 0, name


def test_extended_rejects_points_outside_even_subgroup(ctx, c_oracle):
    """ristretto255 = 2E / E[4]: a curve point with an odd 8-torsion component represents no element.  The validated
    ingestion rejects it; the unchecked form is for trusted representatives and agrees with the validated one on them."""
    import random
    import zkvm_b200 as zk
    from oracle import ristretto255_ref as ref
    rnd = random.Random(11)
    P, L = ref.P, ref.L

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

    t8 = None
    while t8 is None:                                    # a point of exact order 8
        q = smul(L, on_curve_point())
        aff = smul(4, q)
        if aff.X * pow(aff.Z, P - 2, P) % P != 0 or aff.Y * pow(aff.Z, P - 2, P) % P != 1: t8 = q

    def ser(p, lam):
        return b"".join((v * lam % P).to_bytes(32, "little") for v in (p.X, p.Y, p.Z, p.T))

    good_pts = [on_curve_point().double() for _ in range(12)] + [ref.Point.identity(), ref.Point(0, P - 1, 1, 0), smul(2, t8), smul(6, t8)]
    good = [ser(p, rnd.getrandbits(200) + 3) for p in good_pts]
    t = zk.PointTable(ctx).append_extended(b"".join(good))
    tu = zk.PointTable(ctx).append_extended_unchecked(b"".join(good))
    assert len(t) == len(good) and t.compress() == tu.compress()
    assert t.compress() == b"".join(p.encode() for p in good_pts)
    for k in (1, 3, 5, 7):                               # odd multiples of the order-8 point leave 2E
        bad = ser(good_pts[0] + smul(k, t8), rnd.getrandbits(200) + 3)
        tt = zk.PointTable(ctx)
        with pytest.raises(zk.InvalidPoint) as e:
            tt.append_extended(good[1] + good[2] + bad + good[3])
        assert e.value.index == 2 and len(tt) == 0, k
    tz = zk.PointTable(ctx)
    with pytest.raises(zk.InvalidPoint) as e:               # the unchecked form still refuses Z = 0
        tz.append_extended_unchecked(good[0] + b"".join(v.to_bytes(32, "little") for v in (5, 7, 0, 9)) + good[1])
    assert e.value.index == 1 and len(tz) == 0
    # bulk: 3000 valid representatives through the batched-inversion path == the decoder's table
    n = 3000
    comp = make_points(c_oracle, n, 5150)
    blob = b"".join(ser(ref.decode(comp[32 * i:32 * i + 32]), rnd.getrandbits(128) + 2) for i in range(0, n, 10))
    tb = zk.PointTable(ctx).append_extended_unchecked(blob)
    assert tb.compress() == b"".join(comp[32 * i:32 * i + 32] for i in range(0, n, 10))


def test_precomputed_window_tables(ctx, c_oracle):
    """A window-expanded static table gives bit-identical results (single, sliced, device-pointer and batched forms)."""
    import zkvm_b200 as zk
    n = 5000
    gens = make_points(c_oracle, n, 77); sc = rand_scalars(n, 77)
    want = c_oracle.msm(sc, gens, n, threads=4)
    for c in (0, 4, 9, 16, 20):
        tab = zk.PointTable(ctx).append_compressed(gens).precompute(c)
        assert tab.precomputed_window in ((c or 13), (c or 13) - 1)
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == want, c
        off, m = 1234, 2000
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc[:m], tab, offset=off)) == \
            c_oracle.msm(sc[:m], gens[32 * off:32 * (off + m)], m, threads=2), c
        assert tab.compress() == gens                                       # the plain rows are untouched
    sizes = [2500, 0, 1, 777, 5000]
    seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    bs = rand_scalars(int(seg[-1]), 78)
    got = zk.batch_vartime_multiscalar_mul(ctx, bs, tab, seg)
    for k, (a, b) in enumerate(zip(seg, seg[1:])):
        a, b = int(a), int(b)
        assert bytes(got[k]) == c_oracle.msm(bs[a:b], gens[:32 * (b - a)], b - a, threads=2), k
    # adversarial: every scalar equal -> one bucket holds everything
    same = bytes(sc[0]) * n
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, same, tab)) == c_oracle.msm(same, gens, n, threads=4)
    # appending drops the expansion, results stay right
    tab.append_compressed(gens[:64])
    assert tab.precomputed_window == 0
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab, n=n)) == want


def test_block_scale_2e22_negation_and_split(ctx, c_oracle):
    """BASELINE config 5's size on one GPU: 2^22 points.  sum s_i P_i + sum (l - s_i) P_i encodes to the identity, and
    index-range partial sums (the multi-GPU sharding identity) add up to the whole."""
    import zkvm_b200 as zk
    n = 1 << 22
    u = np.random.default_rng(22).integers(0, 256, size=(n // 2, 64), dtype=np.uint8)
    tab = zk.PointTable(ctx, n).append_uniform(u)
    tab.append_compressed(tab.compress())                     # the same 2^21 points again: table of 2^22
    assert len(tab) == n
    sc = rand_scalars(n // 2, 22)
    s_int = [int.from_bytes(bytes(r), "little") % L for r in sc]
    neg = np.frombuffer(b"".join(le32(L - v) for v in s_int), dtype=np.uint8).reshape(-1, 32)
    both = np.concatenate([sc, neg])
    assert zk.RistrettoPoint.vartime_multiscalar_mul(ctx, both, tab).is_identity()
    # positive half only, whole vs 8 shards
    whole = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab, n=n // 2)
    assert not whole.is_identity()
    from zkvm_b200.sharded import shard_range
    parts = b""
    for r in range(8):
        lo, hi = shard_range(n // 2, r, 8)
        parts += bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc[lo:hi], tab, offset=lo))
    assert c_oracle.point_sum(parts, 8) == bytes(whole)


def test_large_2e24_window_independence_and_cancellation(ctx):
    """A single-GPU MSM well past the headline size (2^24 terms, 16.8 M): two window widths agree, and
    sum s_i P_i + sum (l - s_i) P_i over the same points encodes to the identity."""
    import zkvm_b200 as zk
    n = 1 << 24
    half = n // 2
    rng = np.random.default_rng(24)
    tab = zk.PointTable(ctx, half)
    for _ in range(4):                                           # 2^23 points, appended in pieces (growth path)
        tab.append_uniform(rng.integers(0, 256, size=(half // 4, 64), dtype=np.uint8))
    sc = rng.integers(0, 256, size=(half, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0f                                            # < 2^252 < l, so l - s is the exact negation
    s_le = sc.view(np.uint64).reshape(half, 4)
    l_words = np.array([(L >> (64 * k)) & (2**64 - 1) for k in range(4)], dtype=np.uint64)
    neg = np.zeros_like(s_le); borrow = np.zeros(half, dtype=np.uint64)
    for k in range(4):                                           # vectorised 256-bit l - s
        a = np.full(half, l_words[k], dtype=np.uint64); b = s_le[:, k]
        d = a - b - borrow
        borrow = ((a < b) | ((a == b) & (borrow == 1))).astype(np.uint64)
        neg[:, k] = d
    neg8 = neg.view(np.uint8).reshape(half, 32)
    r16 = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
    ctx.set_window(14)
    r14 = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)
    ctx.set_window(0)
    assert bytes(r16) == bytes(r14) and not r16.is_identity()
    rneg = zk.RistrettoPoint.vartime_multiscalar_mul(ctx, neg8, tab)
    # P + (-P) = identity: feed the two 32-byte results back as a 2-term MSM with unit scalars
    assert zk.RistrettoPoint.optional_multiscalar_mul(ctx, le32(1) * 2, bytes(r16) + bytes(rneg)).is_identity()


def test_argument_errors_and_context_reuse(ctx, c_oracle):
    import zkvm_b200 as zk
    pts = make_points(c_oracle, 64, 3); sc = rand_scalars(64, 3)
    with pytest.raises(ValueError):
        zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc[:10], pts)               # length mismatch
    with pytest.raises(ValueError):
        zk.RistrettoPoint.optional_multiscalar_mul(ctx, b"\x00" * 33, pts[:32])      # not a multiple of 32
    with pytest.raises(zk.ZkError):
        ctx.set_window(3)
    with pytest.raises(zk.ZkError):
        ctx.set_window(21)
    tab = zk.PointTable(ctx).append_compressed(pts)
    with pytest.raises(zk.ZkError):
        zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab, offset=1)            # slice past the end
    with pytest.raises(ValueError):
        zk.batch_optional_multiscalar_mul(ctx, sc, pts, [1, 64])                     # segments must start at 0
    with pytest.raises(zk.ZkError):
        zk.batch_optional_multiscalar_mul(ctx, sc, pts, [0, 40, 30, 64])             # not ascending
    # the context is still good after every failure
    assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab)) == c_oracle.msm(sc, pts, 64)
    tab.clear()
    assert len(tab) == 0
    assert zk.RistrettoPoint.vartime_multiscalar_mul(ctx, b"", tab).is_identity()
    # two contexts on one device interleave safely
    c2 = zk.Context(0)
    t2 = zk.PointTable(c2).append_compressed(pts)
    for _ in range(3):
        a = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)
        b = zk.RistrettoPoint.vartime_multiscalar_mul(c2, sc, t2)
        assert bytes(a) == bytes(b)
    t2.close(); c2.close()


# ---------------- several contexts in flight (stream priorities, shared device) ----------------
def test_contexts_in_flight_match_oracle(c_oracle):
    """Six contexts keep MSMs of different sizes in flight on one device from six host threads (the end-to-end shape:
    each context runs its sort / tree / tail on its high-priority stream and its bulk kernels at low priority, and they
    all share the SMs); every result must equal the oracle's on the same bytes, and a context must stay usable."""
    from concurrent.futures import ThreadPoolExecutor
    import zkvm_b200 as zk
    sizes = [1, 37, 1000, 4096, 30000, 70000]
    jobs = []
    for k, n in enumerate(sizes):
        pts = make_points(c_oracle, n, 900 + k)
        sc = rand_scalars(n, 900 + k)
        jobs.append((n, sc, pts, c_oracle.msm(sc, pts, n, threads=4)))
    ctxs = [zk.Context(0) for _ in sizes]
    tabs = [zk.PointTable(ctxs[k], n).append_compressed(jobs[k][2]) for k, (n, *_r) in enumerate(jobs)]

    def worker(k):
        n, sc, pts, want = jobs[k]
        out = []
        for rep in range(6):
            got = zk.RistrettoPoint.optional_multiscalar_mul(ctxs[k], sc, pts) if rep % 2 == 0 else \
                zk.RistrettoPoint.vartime_multiscalar_mul(ctxs[k], sc, tabs[k])
            out.append(got is not None and bytes(got) == want)
        return out

    with ThreadPoolExecutor(max_workers=len(sizes)) as pool:
        results = list(pool.map(worker, range(len(sizes))))
    assert all(all(r) for r in results), results
    # and every context, one after the other, over the same cached table
    n, sc, pts, want = jobs[-1]
    outs = [zk.RistrettoPoint.vartime_multiscalar_mul(cx, sc, tabs[-1]) for cx in ctxs]
    assert all(bytes(o) == want for o in outs)
    for t in tabs: t.close()
    for cx in ctxs: cx.close()


def test_fused_decode_scatter_pinned_sources(ctx, c_oracle, rfc_vectors):
    """Page-locked sources of >= 2^17 compressed points take the fused route (k_decompress_scatter: the decoder scatters
    its own terms' digits, queued after the scan); pageable sources keep the separate scatter.  Same bytes, both routes,
    against the oracle: plain, with a cached static prefix, and with one invalid encoding."""
    import zkvm_b200 as zk
    n = (1 << 17) + 333
    pts = np.frombuffer(make_points(c_oracle, n, 77), dtype=np.uint8).copy()
    sc = rand_scalars(n, 77).reshape(-1).copy()
    want = c_oracle.msm(sc, pts, n, threads=8)
    zk.host_register(pts); zk.host_register(sc)
    try:
        fused = zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)
        assert fused is not None and bytes(fused) == want
        # static prefix from a cached table + the same dynamic suffix
        m = 1000
        pre_pts = make_points(c_oracle, m, 78); pre_sc = rand_scalars(m, 78)
        tab = zk.PointTable(ctx, m).append_compressed(pre_pts)
        got = zk.RistrettoPoint.mixed_multiscalar_mul(ctx, pre_sc, tab, sc, pts)
        all_sc = np.concatenate([pre_sc.reshape(-1), sc]); all_pts = np.concatenate([np.frombuffer(pre_pts, dtype=np.uint8), pts])
        assert got is not None and bytes(got) == c_oracle.msm(all_sc, all_pts, m + n, threads=8)
        # one bad encoding deep inside the dynamic part
        bad = pts.copy(); zk.host_register(bad)
        try:
            k = (1 << 16) + 17
            bad[32 * k:32 * k + 32] = np.frombuffer(H(rfc_vectors["bad_encodings"]["non_canonical"][0]), dtype=np.uint8)
            assert zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, bad) is None
        finally:
            zk.host_unregister(bad)
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want      # the context stays usable
    finally:
        zk.host_unregister(pts); zk.host_unregister(sc)
    # the pageable route (staging ring, separate scatter) on the same bytes
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, np.array(sc), np.array(pts))) == want
