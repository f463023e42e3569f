"""The single-process multi-GPU entry points of the C ABI (zk_mgpu_*) against the oracle.  Uses every GPU of the box
(1 on the driver's test box: the same code path with one shard), both gather modes, pageable and pinned sources, and a
C++ program through the header-only mirror at the block-scale sizes BASELINE.json config 5 names."""
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _g():
    g = min(torch.cuda.device_count(), 8)
    assert g >= 1
    return g


def _inputs(c_oracle, n, seed):
    rng = np.random.default_rng(seed)
    pts = c_oracle.from_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8), n)
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    return sc, pts


def _nccl_ok(mg):
    import zkvm_b200 as zk
    try:
        mg.set_gather("nccl")
        return True
    except zk.ZkError:
        return False


@pytest.mark.parametrize("n", [0, 1, 7, 5000, (1 << 16) + 3])
def test_mgpu_compressed_matches_oracle(c_oracle, n):
    import zkvm_b200 as zk
    sc, pts = _inputs(c_oracle, n, 31 + n)
    want = c_oracle.msm(sc, pts, n, threads=4)
    mg = zk.MultiGpu(g=_g())
    assert bytes(mg.optional_multiscalar_mul(sc, pts)) == want
    mg.set_staging(2)                                   # every upload through the pinned ring (the pageable-source path)
    assert bytes(mg.optional_multiscalar_mul(sc, pts)) == want
    mg.set_staging(0)
    if _nccl_ok(mg):
        assert bytes(mg.optional_multiscalar_mul(sc, pts)) == want
    assert mg.launch_count > 0 or n == 0
    mg.close()


def test_mgpu_invalid_point_is_none(c_oracle, rfc_vectors):
    import zkvm_b200 as zk
    n = 4001
    sc, pts = _inputs(c_oracle, n, 77)
    mg = zk.MultiGpu(g=_g())
    for idx in (0, n // 2, n - 1):
        bad = bytearray(pts); bad[32 * idx:32 * idx + 32] = bytes.fromhex(rfc_vectors["bad_encodings"]["negative_s"][1])
        assert mg.optional_multiscalar_mul(sc, bytes(bad)) is None
        assert bytes(mg.optional_multiscalar_mul(sc, pts)) == c_oracle.msm(sc, pts, n, threads=4)     # the handle stays usable
        t = zk.MultiGpuTable(mg)
        t.append_compressed(pts[:32 * 100])
        with pytest.raises(zk.InvalidPoint) as e:
            t.append_compressed(bytes(bad))
        assert e.value.index == idx and len(t) == 100                                                # all-or-nothing
        assert bytes(mg.vartime_multiscalar_mul(sc[:100], t)) == c_oracle.msm(sc[:100], pts[:32 * 100], 100)
        t.close()
    mg.close()


def test_mgpu_sharded_table_slices(c_oracle):
    """Several appends, each cut into g ranges; any global slice maps to one contiguous run of rows per device."""
    import zkvm_b200 as zk
    sizes = [1000, 1, 0, 4097, 333]
    n = sum(sizes)
    sc, pts = _inputs(c_oracle, n, 99)
    for gather in ("peer", "nccl"):
        mg = zk.MultiGpu(g=_g())
        if gather == "nccl" and not _nccl_ok(mg):
            mg.close(); continue
        t = zk.MultiGpuTable(mg)
        pos = 0
        for k in sizes:
            t.append_compressed(pts[32 * pos:32 * (pos + k)]); pos += k
        assert len(t) == n
        rng = np.random.default_rng(5)
        cases = [(0, n), (0, 0), (n, 0), (999, 3), (1000, 1), (1001, 4097), (500, 4000)]
        for _ in range(6):
            a, b = sorted(int(v) for v in rng.integers(0, n + 1, size=2))
            cases.append((a, b - a))
        for off, m in cases:
            got = mg.vartime_multiscalar_mul(sc[off:off + m], t, offset=off)
            assert bytes(got) == c_oracle.msm(sc[off:off + m], pts[32 * off:32 * (off + m)], m, threads=2), (off, m)
        with pytest.raises(zk.ZkError):
            mg.vartime_multiscalar_mul(sc[:10], t, offset=n - 5)
        # mixed form: a slice of the sharded cache + dynamic compressed points (bulletproofs' verification shape)
        for off, m, nd in ((0, n, 0), (700, 3000, 77), (0, 0, 500), (n - 1, 1, 1), (1000, 1, 4099)):
            dsc, dpt = _inputs(c_oracle, nd, 1000 + nd)
            all_s = np.concatenate([sc[off:off + m].reshape(-1), dsc.reshape(-1)])
            all_p = pts[32 * off:32 * (off + m)] + dpt
            got = mg.mixed_multiscalar_mul(sc[off:off + m], t, dsc, dpt, offset=off)
            assert bytes(got) == c_oracle.msm(all_s, all_p, m + nd, threads=2), (off, m, nd)
        assert mg.mixed_multiscalar_mul(sc[:5], t, sc[:1], bytes([1]) + bytes(31)) is None       # s = 1 is not a valid encoding
        u = np.random.default_rng(8).integers(0, 256, size=(257, 64), dtype=np.uint8)
        t.append_uniform(u)
        up = c_oracle.from_uniform(u, 257)
        s2 = np.random.default_rng(9).integers(0, 256, size=(257, 32), dtype=np.uint8)
        assert bytes(mg.vartime_multiscalar_mul(s2, t, offset=n)) == c_oracle.msm(s2, up, 257, threads=2)
        t.close(); mg.close()


def test_single_gpu_staging_ring(ctx, c_oracle):
    """Pageable sources go through the ctx's pinned ring; pinned (registered) sources do not; same bytes either way."""
    import zkvm_b200 as zk
    n = 40000                                           # 1.25 MiB per array: above the 256 KiB staging threshold
    sc, pts = _inputs(c_oracle, n, 1234)
    pts = np.frombuffer(pts, dtype=np.uint8).copy()
    want = c_oracle.msm(sc, pts, n, threads=4)
    b0 = ctx.staged_bytes
    assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want
    assert ctx.staged_bytes - b0 == 64 * n              # both arrays were pageable numpy memory
    zk.host_register(sc); zk.host_register(pts)
    try:
        b1 = ctx.staged_bytes
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want
        assert ctx.staged_bytes == b1                   # page-locked now: direct DMA
    finally:
        zk.host_unregister(sc); zk.host_unregister(pts)
    ctx.set_staging(1)
    try:
        assert bytes(zk.RistrettoPoint.optional_multiscalar_mul(ctx, sc, pts)) == want
    finally:
        ctx.set_staging(0)


def test_sharded_module_single_process(c_oracle):
    from zkvm_b200.sharded import msm_single_process
    n = 3000
    sc, pts = _inputs(c_oracle, n, 42)
    assert msm_single_process(sc, pts, g=_g()) == c_oracle.msm(sc, pts, n, threads=2)


CPP = r'''
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "zkvm_b200/cpp/ristretto_msm.hpp"
using namespace zkvm_b200;
// argv: g, gather (0 peer / 1 nccl), file with n (u64), n*32 scalars, n*32 compressed points.
// Prints the encoding from the compressed path and from a sharded table filled by two appends.
int main(int, char** argv) {
    int g = atoi(argv[1]), gather = atoi(argv[2]);
    FILE* f = fopen(argv[3], "rb"); unsigned long long n; if (fread(&n, 8, 1, f) != 1) return 2;
    std::vector<Scalar> s(n); std::vector<CompressedRistretto> p(n);
    if (fread(s.data(), 32, n, f) != n || fread(p.data(), 32, n, f) != n) return 2;
    fclose(f);
    std::vector<int> dev(g); for (int i = 0; i < g; i++) dev[i] = i;
    MultiGpu mg(dev, gather ? MultiGpu::Gather::Nccl : MultiGpu::Gather::Peer);
    auto r1 = mg.optional_multiscalar_mul(s.data(), p.data(), n);
    if (!r1) { printf("none\n"); return 0; }
    for (auto b : *r1) printf("%02x", b); printf("\n");
    MultiGpuTable t(mg);
    if (t.append_compressed(p.data(), n / 3)) return 3;
    if (t.append_compressed(p.data() + n / 3, n - n / 3)) return 3;
    auto r2 = t.vartime_multiscalar_mul(s.data(), 0, n);
    for (auto b : r2) printf("%02x", b); printf("\n");
    return 0;
}
'''


@pytest.mark.parametrize("log2n", [16, 22, 23])
def test_cpp_mgpu_block_scale(tmp_path, ctx, c_oracle, log2n):
    """A C++ host program drives every GPU of the box through the header-only mirror at config 5's sizes (2^22, 2^23)
    and must print the oracle's bytes.  Points come from the GPU's hash-to-group (itself pinned to the oracle and the
    RFC vectors in test_gpu_parity.py): the CPU oracle's from_uniform is single-threaded and would take minutes."""
    import zkvm_b200 as zk
    src = tmp_path / "m.cpp"; src.write_text(CPP)
    exe = tmp_path / "m"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", ROOT, str(src), "-o", str(exe), "-L", os.path.join(ROOT, "zkvm_b200"),
                           "-l:libzkmsm.so", f"-Wl,-rpath,{os.path.join(ROOT, 'zkvm_b200')}"])
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    tab = zk.PointTable(ctx, n).append_uniform(rng.integers(0, 256, size=(n, 64), dtype=np.uint8))
    pts = tab.compress(); tab.close()
    sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    blob = tmp_path / "in.bin"
    with open(blob, "wb") as f:
        f.write(np.uint64(n).tobytes()); f.write(sc.tobytes()); f.write(pts)
    want = c_oracle.msm(sc, pts, n, threads=min(32, os.cpu_count() or 1)).hex()
    probe = zk.MultiGpu(g=_g())
    modes = [0, 1] if _nccl_ok(probe) else [0]          # without a loadable libnccl the peer-copy gather is the only mode
    probe.close()
    for gather in modes:
        out = subprocess.check_output([str(exe), str(_g()), str(gather), str(blob)], text=True).split()
        assert out[-2:] == [want, want], (log2n, gather, out)      # NCCL may print a version banner first
