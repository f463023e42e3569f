"""(1) The C++ host mirror end to end on the GPU: a small C++ program through ristretto_msm.hpp, checked against the
oracle.  (2) A bounded fuzz over sizes, window widths, scalar shapes and entry points."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CPP = r'''
#include <cstdio>
#include <cstring>
#include <vector>
#include "zkvm_b200/cpp/ristretto_msm.hpp"
using namespace zkvm_b200;
// argv[1]: file with n (u64), n*32 scalars, n*32 compressed points.  Prints hex of both entry points' results.
int main(int, char** argv) {
    FILE* f = fopen(argv[1], "rb"); unsigned long long n; if (fread(&n, 8, 1, f) != 1) return 2;
    std::vector<Scalar> s(n); std::vector<CompressedRistretto> p(n);
    if (fread(s.data(), 32, n, f) != n || fread(p.data(), 32, n, f) != n) return 2;
    fclose(f);
    Context ctx(0);
    auto r1 = RistrettoPoint::optional_multiscalar_mul(ctx, s, p);
    if (!r1) { printf("none\n"); return 0; }
    PointTable t(ctx);
    if (t.append_compressed(p.data(), n)) return 3;
    auto r2 = RistrettoPoint::vartime_multiscalar_mul(ctx, s, t);
    for (auto b : *r1) printf("%02x", b); printf("\n");
    for (auto b : r2) printf("%02x", b); printf("\n");
    printf("%d\n", (int)RistrettoPoint::is_identity(r2));
    // window-expanded table, mixed form, and a 3-MSM batch whose pieces cover everything
    t.precompute(0);
    auto r3 = RistrettoPoint::vartime_multiscalar_mul(ctx, s, t);
    for (auto b : r3) printf("%02x", b); printf("\n");
    std::vector<Scalar> s_st(s.begin(), s.begin() + n / 2), s_dy(s.begin() + n / 2, s.end());
    std::vector<CompressedRistretto> p_dy(p.begin() + n / 2, p.end());
    auto r4 = RistrettoPoint::mixed_multiscalar_mul(ctx, s_st, t, 0, s_dy, p_dy);
    for (auto b : *r4) printf("%02x", b); printf("\n");
    std::vector<uint64_t> seg = {0, n / 3, n / 3, n};
    auto rb = RistrettoPoint::batch_optional_multiscalar_mul(ctx, s, p, seg);
    printf("%d %d\n", (int)rb.size(), (int)RistrettoPoint::is_identity(*rb[1]));
    return 0;
}
'''


def test_cpp_mirror_end_to_end(tmp_path, c_oracle, rfc_vectors):
    src = tmp_path / "t.cpp"; src.write_text(CPP)
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", ROOT, str(src), "-o", str(exe), "-L", os.path.join(ROOT, "zkvm_b200"),
                           "-l:libzkmsm.so", f"-Wl,-rpath,{os.path.join(ROOT, 'zkvm_b200')}"])
    n = 777
    rng = np.random.default_rng(8)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    blob = tmp_path / "in.bin"
    blob.write_bytes(np.uint64(n).tobytes() + sc.tobytes() + pts)
    out = subprocess.check_output([str(exe), str(blob)], text=True).split()
    want = c_oracle.msm(sc, pts, n).hex()
    assert out == [want, want, "0", want, want, "3", "1"]
    bad = bytearray(pts); bad[32 * 5:32 * 6] = bytes.fromhex(rfc_vectors["bad_encodings"]["negative_s"][2])
    blob.write_bytes(np.uint64(n).tobytes() + sc.tobytes() + bytes(bad))
    assert subprocess.check_output([str(exe), str(blob)], text=True).split() == ["none"]


def test_fuzz_sizes_windows_shapes(ctx, c_oracle):
    import zkvm_b200 as zk
    L = zk.GROUP_ORDER
    rng = np.random.default_rng(20260)
    pool_n = 6000
    pool = c_oracle.from_uniform(rng.integers(0, 256, size=(pool_n, 64), dtype=np.uint8), pool_n)
    tab = zk.PointTable(ctx).append_compressed(pool)
    pre = zk.PointTable(ctx).append_compressed(pool).precompute(int(rng.integers(4, 21)))
    for trial in range(40):
        n = int(rng.choice([1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 255, 256, 257, 1000, 4095, 4096, 4097, int(rng.integers(1, pool_n))]))
        off = int(rng.integers(0, pool_n - n + 1))
        shape = trial % 5
        if shape == 0: sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)                       # any 256-bit string
        elif shape == 1: sc = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in rng.integers(0, 4, size=n)), dtype=np.uint8).reshape(n, 32)  # tiny
        elif shape == 2: sc = np.tile(rng.integers(0, 256, size=(1, 32), dtype=np.uint8), (n, 1))     # all equal
        elif shape == 3: sc = np.frombuffer(b"".join(((L - 1 - int(v)) % 2**256).to_bytes(32, "little") for v in rng.integers(0, 3, size=n)), dtype=np.uint8).reshape(n, 32)  # near l
        else:                                                                                        # one nonzero window each
            sc = np.frombuffer(b"".join((int(rng.integers(1, 1 << 16)) << int(rng.integers(0, 237))).to_bytes(32, "little") for _ in range(n)), dtype=np.uint8).reshape(n, 32)
        pts = pool[32 * off:32 * (off + n)]
        want = c_oracle.msm(sc, pts, n, threads=2)
        ctx.set_window(int(rng.choice([0, 0, 4, 5, 7, 9, 12, 14, 16])))
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want, (trial, n, shape)
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, tab, offset=off)) == want, (trial, n, shape)
        assert bytes(zk.RistrettoPoint.vartime_multiscalar_mul(ctx, sc, pre, offset=off)) == want, (trial, n, shape)
        k = int(rng.integers(0, n + 1))
        got = zk.RistrettoPoint.mixed_multiscalar_mul(ctx, sc[:k], tab, sc[k:], pts[32 * k:], offset=off)
        assert bytes(got) == want, (trial, n, shape, k)
    ctx.set_window(0)
