"""ctypes loader for oracle/libmsm_oracle.so (the C restatement, msm_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmsm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "msm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        # x86-64-v3 (AVX2, BMI2, ADX-era servers): the .so is built here and runs on the GPU box, so no -march=native
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-o", _SO, src, "-lpthread"])
    return _SO


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_SO)
            _lib.oracle_is_valid(bytes(32))
        except OSError:
            build(force=True)
            _lib = C.CDLL(_SO)
        _lib.oracle_ge_size.restype = C.c_size_t
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if hasattr(a, "ctypes") else a


def msm(scalars, points, n: int, threads: int = 1):
    """decode + MSM + encode.  Returns 32 bytes, or None if any encoding is invalid."""
    out = C.create_string_buffer(32)
    rc = load().oracle_msm(_p(scalars), _p(points), C.c_size_t(n), threads, out)
    return None if rc else out.raw


def decompress(points, n: int, threads: int = 1):
    """Returns (opaque decompressed buffer, bad_index or None)."""
    lib = load()
    buf = C.create_string_buffer(max(1, n) * lib.oracle_ge_size())
    bad = C.c_size_t(0)
    rc = lib.oracle_decompress(_p(points), C.c_size_t(n), threads, buf, C.byref(bad))
    return buf, (bad.value if rc else None)


def msm_decompressed(scalars, ge_buf, n: int, threads: int = 1) -> bytes:
    out = C.create_string_buffer(32)
    load().oracle_msm_decompressed(_p(scalars), ge_buf, C.c_size_t(n), threads, out)
    return out.raw


def from_uniform(bytes64, n: int) -> bytes:
    out = C.create_string_buffer(max(1, 32 * n))
    load().oracle_from_uniform(_p(bytes64), C.c_size_t(n), out)
    return out.raw[: 32 * n]


def scalarmult(s32: bytes, p32: bytes):
    out = C.create_string_buffer(32)
    return None if load().oracle_scalarmult(s32, p32, out) else out.raw


def point_sum(points, n: int):
    out = C.create_string_buffer(32)
    return None if load().oracle_sum(_p(points), C.c_size_t(n), out) else out.raw


def is_valid(p32: bytes) -> bool:
    return bool(load().oracle_is_valid(p32))


def set_vector(on: bool) -> bool:
    """Select the 4-lane AVX-512 IFMA bucket accumulation (dalek's vector-backend shape) for the Pippenger MSMs of this
    process.  Returns whether it is in use (False on a CPU without IFMA: the scalar radix-2^51 code keeps running)."""
    return bool(load().oracle_set_vector(1 if on else 0))
