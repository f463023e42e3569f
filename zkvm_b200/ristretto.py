"""Host-side mirror of the reference-facing API for the MSM hot path.

The names follow curve25519-dalek's public items for this path -- `CompressedRistretto`,
`RistrettoPoint`, `Scalar`, and the `VartimeMultiscalarMul` trait methods
`vartime_multiscalar_mul` / `optional_multiscalar_mul` -- as recalled from public knowledge
(unverified: the upstream source is not mounted and /root/reference has
no code to cite).  Behaviour at the byte level follows RFC 9496.  Everything here forwards to
the CUDA library through the C ABI in include/zkmsm.h; nothing is computed on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

from . import _lib
from ._lib import ZkError

IDENTITY_BYTES = bytes(32)
GROUP_ORDER = 2**252 + 27742317777372353535851937790883648493


def _as_buf(x) -> tuple[object, int]:
    """bytes-like / numpy uint8 array -> (ctypes-compatible pointer holder, nbytes)."""
    if isinstance(x, (bytes, bytearray)):
        return (C.c_char * len(x)).from_buffer_copy(x) if isinstance(x, bytes) else (C.c_char * len(x)).from_buffer(x), len(x)
    try:
        import numpy as np
        if isinstance(x, np.ndarray):
            a = np.ascontiguousarray(x, dtype=np.uint8)
            return a, a.nbytes
    except ImportError:      # pragma: no cover
        pass
    mv = memoryview(x).cast("B")
    return (C.c_char * len(mv)).from_buffer_copy(mv), len(mv)


def _ptr(holder) -> C.c_void_p:
    if hasattr(holder, "ctypes"):
        return C.c_void_p(holder.ctypes.data)
    return C.cast(holder, C.c_void_p)


def _join32(items) -> tuple[object, int]:
    """Sequence of 32-byte strings, Scalars, CompressedRistrettos, one bytes blob or a (n,32) uint8 array."""
    if isinstance(items, (bytes, bytearray)) or hasattr(items, "ctypes"):
        h, nb = _as_buf(items)
    else:
        blob = b"".join(bytes(i) for i in items)
        h, nb = _as_buf(blob)
    if nb % 32:
        raise ValueError("expected a whole number of 32-byte items")
    return h, nb // 32


class Scalar:
    """32-byte little-endian scalar (dalek `Scalar`): any 256-bit pattern; used modulo the group order."""
    __slots__ = ("_b",)

    def __init__(self, b: bytes):
        if len(b) != 32:
            raise ValueError("Scalar needs 32 bytes")
        self._b = bytes(b)

    @classmethod
    def from_int(cls, v: int) -> "Scalar":
        return cls((v % GROUP_ORDER).to_bytes(32, "little"))

    def __bytes__(self) -> bytes:
        return self._b

    def as_bytes(self) -> bytes:
        return self._b


class CompressedRistretto:
    """32-byte ristretto255 encoding (dalek `CompressedRistretto`)."""
    __slots__ = ("_b",)

    def __init__(self, b: bytes):
        if len(b) != 32:
            raise ValueError("CompressedRistretto needs 32 bytes")
        self._b = bytes(b)

    def __bytes__(self) -> bytes:
        return self._b

    def as_bytes(self) -> bytes:
        return self._b

    def __eq__(self, o) -> bool:
        return isinstance(o, CompressedRistretto) and o._b == self._b

    def __hash__(self):
        return hash(self._b)

    def is_identity(self) -> bool:
        return self._b == IDENTITY_BYTES

    def decompress(self, ctx: "Context") -> Optional["PointTable"]:
        """RFC 9496 4.3.1 on the device; None when the encoding is rejected (dalek returns Option)."""
        t = PointTable(ctx, 1)
        try:
            t.append_compressed(self._b)
        except InvalidPoint:
            return None
        return t


class InvalidPoint(ZkError):
    def __init__(self, index: Optional[int]):
        self.index = index
        ZkError.__init__(self, _lib.ZK_ERR_INVALID_POINT, f"invalid ristretto255 encoding at index {index}")


class Context:
    """One per thread/process and device: stream + reusable HBM workspace (zk_ctx)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.zk_ctx_create(device, C.byref(h))
        if rc != 0:
            raise ZkError(rc, f"zk_ctx_create(device={device}) failed: {self._lib.zk_status_str(rc).decode()} -- "
                              "a CUDA device is required, there is no CPU fallback")
        self._h = h
        self.device = device

    def _check(self, rc: int, bad_index: Optional[int] = None):
        if rc == 0:
            return
        if rc == _lib.ZK_ERR_INVALID_POINT:
            raise InvalidPoint(bad_index)
        raise ZkError(rc, f"{self._lib.zk_status_str(rc).decode()}: {self._lib.zk_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.zk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self._lib.zk_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        return self._lib.zk_ctx_stream(self._h) or 0

    def set_window(self, c: int):
        self._check(self._lib.zk_ctx_set_window(self._h, c))

    def set_profiling(self, on: bool):
        self._check(self._lib.zk_ctx_set_profiling(self._h, 1 if on else 0))

    def last_phase_ms(self) -> list[float]:
        arr = (C.c_float * 4)()
        self._check(self._lib.zk_ctx_last_phase_ms(self._h, C.byref(arr)))
        return list(arr)

    @property
    def launch_count(self) -> int:
        return int(self._lib.zk_ctx_launch_count(self._h))

    def set_priority(self, high: bool):
        """Recreate the context's streams with the highest (or default) stream priority; call while idle."""
        self._check(self._lib.zk_ctx_set_priority(self._h, 1 if high else 0))

    def sum_compressed(self, points) -> Optional["CompressedRistretto"]:
        """Sum of up to 1024 compressed points in one small launch (`iter.sum()` over decompressed points, compressed);
        None if an encoding is invalid."""
        h, g = _join32(points)
        out = C.create_string_buffer(32)
        rc = self._lib.zk_sum_compressed(self._h, _ptr(h), g, out)
        if rc == _lib.ZK_ERR_INVALID_POINT:
            return None
        self._check(rc)
        return CompressedRistretto(out.raw)

    def set_wait(self, mode: int):
        """0 = spin while waiting for the device (default), 1 = block (no CPU while waiting, ~0.3 ms more latency)."""
        self._check(self._lib.zk_ctx_set_wait(self._h, mode))

    def set_staging(self, mode: int):
        """0 = auto (pinned ring for pageable sources), 1 = never stage, 2 = always stage."""
        self._check(self._lib.zk_ctx_set_staging(self._h, mode))

    @property
    def staged_bytes(self) -> int:
        return int(self._lib.zk_ctx_staged_bytes(self._h))

    def bench_int_pipe(self, kind: int) -> float:
        v = C.c_double()
        self._check(self._lib.zk_bench_int_pipe(self._h, kind, C.byref(v)))
        return v.value

    # ---- device-pointer forms (plumbing for torch tensors) ----
    def msm_table_dev(self, scalars_dev_ptr: int, table: "PointTable", offset: int, n: int, out_ext_dev_ptr: int):
        self._check(self._lib.zk_msm_table_dev(self._h, C.c_void_p(scalars_dev_ptr), table._h, offset, n,
                                               C.c_void_p(out_ext_dev_ptr)))

    def ext_sum_compress_dev(self, ext_dev_ptr: int, g: int) -> CompressedRistretto:
        out = C.create_string_buffer(32)
        self._check(self._lib.zk_ext_sum_compress_dev(self._h, C.c_void_p(ext_dev_ptr), g, out))
        return CompressedRistretto(out.raw)


class PointTable:
    """Device-resident cache of decompressed points (zk_table): the GPU-side `Vec<RistrettoPoint>`."""

    def __init__(self, ctx: Context, capacity: int = 0):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(ctx._lib.zk_table_create(ctx._h, capacity, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.zk_table_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.ctx._lib.zk_table_len(self._h))

    def clear(self):
        self.ctx._lib.zk_table_clear(self._h)

    def append_compressed(self, points) -> "PointTable":
        h, n = _join32(points)
        bad = C.c_size_t(0)
        rc = self.ctx._lib.zk_table_append_compressed(self.ctx._h, self._h, _ptr(h), n, C.byref(bad))
        self.ctx._check(rc, bad.value)
        return self

    def append_compressed_dev(self, dev_ptr: int, n: int) -> "PointTable":
        bad = C.c_size_t(0)
        rc = self.ctx._lib.zk_table_append_compressed_dev(self.ctx._h, self._h, C.c_void_p(dev_ptr), n, C.byref(bad))
        self.ctx._check(rc, bad.value)
        return self

    def precompute(self, c: int = 0) -> "PointTable":
        """Expand a static generator table into per-window multiples (no doublings, shared buckets) -- the device-side
        counterpart of dalek's VartimePrecomputedMultiscalarMul.  W x the memory; run once per generator set."""
        self.ctx._check(self.ctx._lib.zk_table_precompute(self.ctx._h, self._h, c))
        return self

    @property
    def precomputed_window(self) -> int:
        return int(self.ctx._lib.zk_table_precomputed_window(self._h))

    def append_extended(self, ext128) -> "PointTable":
        """Already-decompressed points as X, Y, Z, T (4 x 32-byte canonical little-endian field elements each)."""
        h, nb = _as_buf(ext128 if not isinstance(ext128, (list, tuple)) else b"".join(ext128))
        if nb % 128:
            raise ValueError("expected a whole number of 128-byte extended points")
        bad = C.c_size_t(0)
        rc = self.ctx._lib.zk_table_append_extended(self.ctx._h, self._h, _ptr(h), nb // 128, C.byref(bad))
        self.ctx._check(rc, bad.value)
        return self

    def append_extended_unchecked(self, ext128) -> "PointTable":
        """Same input as append_extended for points KNOWN to be valid representatives (e.g. values of a RistrettoPoint):
        only Z = 0 is rejected; normalisation uses a batched inversion (about 7x faster than the validated form)."""
        h, nb = _as_buf(ext128 if not isinstance(ext128, (list, tuple)) else b"".join(ext128))
        if nb % 128:
            raise ValueError("expected a whole number of 128-byte extended points")
        bad = C.c_size_t(0)
        rc = self.ctx._lib.zk_table_append_extended_unchecked(self.ctx._h, self._h, _ptr(h), nb // 128, C.byref(bad))
        self.ctx._check(rc, bad.value)
        return self

    def append_uniform(self, bytes64) -> "PointTable":
        """RistrettoPoint::from_uniform_bytes over a batch of 64-byte strings."""
        h, nb = _as_buf(bytes64 if not isinstance(bytes64, (list, tuple)) else b"".join(bytes64))
        if nb % 64:
            raise ValueError("expected a whole number of 64-byte strings")
        self.ctx._check(self.ctx._lib.zk_table_append_uniform(self.ctx._h, self._h, _ptr(h), nb // 64))
        return self

    def append_uniform_dev(self, dev_ptr: int, n: int) -> "PointTable":
        self.ctx._check(self.ctx._lib.zk_table_append_uniform_dev(self.ctx._h, self._h, C.c_void_p(dev_ptr), n))
        return self

    def compress(self, offset: int = 0, n: Optional[int] = None) -> bytes:
        n = len(self) - offset if n is None else n
        out = C.create_string_buffer(max(32 * n, 1))
        self.ctx._check(self.ctx._lib.zk_table_compress(self.ctx._h, self._h, offset, n, out))
        return out.raw[: 32 * n]

    def compress_dev(self, out_dev_ptr: int, offset: int = 0, n: Optional[int] = None):
        n = len(self) - offset if n is None else n
        self.ctx._check(self.ctx._lib.zk_table_compress_dev(self.ctx._h, self._h, offset, n, C.c_void_p(out_dev_ptr)))


class RistrettoPoint:
    """Namespace for the `VartimeMultiscalarMul` entry points (results are returned compressed)."""

    @staticmethod
    def vartime_multiscalar_mul(ctx: Context, scalars, points: "PointTable", offset: int = 0,
                                n: Optional[int] = None) -> CompressedRistretto:
        """sum scalars[i] * points[offset+i] over cached (already decompressed) points."""
        h, ns = _join32(scalars)
        n = ns if n is None else n
        if ns != n:
            raise ValueError("scalars/points length mismatch")
        out = C.create_string_buffer(32)
        ctx._check(ctx._lib.zk_msm_vartime_table(ctx._h, _ptr(h), points._h, offset, n, out))
        return CompressedRistretto(out.raw)

    @staticmethod
    def optional_multiscalar_mul(ctx: Context, scalars, points) -> Optional[CompressedRistretto]:
        """sum scalars[i] * decompress(points[i]); None if any encoding is invalid
        (the contract of dalek's optional_multiscalar_mul over `points.map(|p| p.decompress())`)."""
        hs, ns = _join32(scalars)
        hp, np_ = _join32(points)
        if ns != np_:
            raise ValueError("scalars/points length mismatch")
        out = C.create_string_buffer(32)
        rc = ctx._lib.zk_msm_vartime(ctx._h, _ptr(hs), _ptr(hp), ns, out)
        if rc == _lib.ZK_ERR_INVALID_POINT:
            return None
        ctx._check(rc)
        return CompressedRistretto(out.raw)

    @staticmethod
    def mixed_multiscalar_mul(ctx: Context, static_scalars, table: Optional["PointTable"], dyn_scalars, dyn_points,
                              offset: int = 0) -> Optional[CompressedRistretto]:
        """Cached generators first, then the proof's own compressed points (bulletproofs' verification shape)."""
        hs, ns = _join32(static_scalars)
        hd, nd = _join32(dyn_scalars)
        hp, np_ = _join32(dyn_points)
        if nd != np_:
            raise ValueError("dynamic scalars/points length mismatch")
        out = C.create_string_buffer(32)
        rc = ctx._lib.zk_msm_vartime_mixed(ctx._h, _ptr(hs), table._h if table is not None else None, offset, ns,
                                           _ptr(hd), _ptr(hp), nd, out)
        if rc == _lib.ZK_ERR_INVALID_POINT:
            return None
        ctx._check(rc)
        return CompressedRistretto(out.raw)


def _seg_array(seg_offsets):
    import numpy as np
    seg = np.ascontiguousarray(seg_offsets, dtype=np.uint64)
    if seg.ndim != 1 or len(seg) < 2 or seg[0] != 0:
        raise ValueError("seg_offsets needs m+1 ascending entries starting at 0")
    return seg


def batch_optional_multiscalar_mul(ctx: Context, scalars, points, seg_offsets) -> list:
    """m independent MSMs in one device pass: MSM k covers terms seg_offsets[k]:seg_offsets[k+1] of the concatenated
    scalars / compressed points.  Returns a list of CompressedRistretto, with None where that MSM had an invalid
    encoding (each proof keeps its own verdict)."""
    hs, ns = _join32(scalars)
    hp, np_ = _join32(points)
    seg = _seg_array(seg_offsets)
    m = len(seg) - 1
    if ns != np_ or ns != int(seg[-1]):
        raise ValueError("scalars/points/segments length mismatch")
    out = C.create_string_buffer(32 * m)
    valid = C.create_string_buffer(m)
    rc = ctx._lib.zk_msm_vartime_batch(ctx._h, _ptr(hs), _ptr(hp), _ptr(seg), m, out, valid)
    if rc not in (0, _lib.ZK_ERR_INVALID_POINT):
        ctx._check(rc)
    return [CompressedRistretto(out.raw[32 * k:32 * k + 32]) if valid.raw[k] else None for k in range(m)]


def batch_vartime_multiscalar_mul(ctx: Context, scalars, table: "PointTable", seg_offsets, offset: int = 0) -> list:
    """m MSMs over the SAME cached points: MSM k = sum_j scalars[seg[k] + j] * table[offset + j]."""
    hs, ns = _join32(scalars)
    seg = _seg_array(seg_offsets)
    m = len(seg) - 1
    if ns != int(seg[-1]):
        raise ValueError("scalars/segments length mismatch")
    out = C.create_string_buffer(32 * m)
    ctx._check(ctx._lib.zk_msm_vartime_table_batch(ctx._h, _ptr(hs), table._h, offset, _ptr(seg), m, out))
    return [CompressedRistretto(out.raw[32 * k:32 * k + 32]) for k in range(m)]


def pick_window(n: int) -> int:
    return int(_lib.load().zk_pick_window(n))


def host_register(arr) -> None:
    """Page-lock a long-lived numpy buffer so that uploads from it are direct asynchronous DMA (zk_host_register)."""
    rc = _lib.load().zk_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes)
    if rc != 0:
        raise ZkError(rc, "zk_host_register failed")


def host_unregister(arr) -> None:
    rc = _lib.load().zk_host_unregister(C.c_void_p(arr.ctypes.data))
    if rc != 0:
        raise ZkError(rc, "zk_host_unregister failed")


class MultiGpuTable:
    """Point cache sharded by index range over the devices of a MultiGpu (zk_mgpu_table)."""

    def __init__(self, mg: "MultiGpu", capacity: int = 0):
        self.mg = mg
        h = C.c_void_p()
        mg._check(mg._lib.zk_mgpu_table_create(mg._h, capacity, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.mg, "_h", None):
            self.mg._lib.zk_mgpu_table_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self.mg._lib.zk_mgpu_table_len(self._h))

    def clear(self):
        self.mg._lib.zk_mgpu_table_clear(self._h)

    def append_compressed(self, points) -> "MultiGpuTable":
        h, n = _join32(points)
        bad = C.c_size_t(0)
        self.mg._check(self.mg._lib.zk_mgpu_table_append_compressed(self._h, _ptr(h), n, C.byref(bad)), bad.value)
        return self

    def append_uniform(self, bytes64) -> "MultiGpuTable":
        h, nb = _as_buf(bytes64 if not isinstance(bytes64, (list, tuple)) else b"".join(bytes64))
        if nb % 64:
            raise ValueError("expected a whole number of 64-byte strings")
        self.mg._check(self.mg._lib.zk_mgpu_table_append_uniform(self._h, _ptr(h), nb // 64))
        return self


class MultiGpu:
    """Several GPUs of one box behind one call (zk_mgpu): point-range shards, one gather of 128-byte partials."""

    def __init__(self, devices: Optional[Sequence[int]] = None, g: Optional[int] = None, gather: str = "peer"):
        self._lib = _lib.load()
        if devices is None:
            if g is None:
                raise ValueError("give devices or g")
            arr = None
        else:
            g = len(devices)
            arr = (C.c_int * g)(*devices)
        h = C.c_void_p()
        rc = self._lib.zk_mgpu_create(arr, g, C.byref(h))
        if rc != 0:
            raise ZkError(rc, f"zk_mgpu_create(g={g}) failed: {self._lib.zk_status_str(rc).decode()} -- CUDA devices are "
                              "required, there is no CPU fallback")
        self._h = h
        self.g = g
        if gather != "peer":
            self.set_gather(gather)

    def _check(self, rc: int, bad_index: Optional[int] = None):
        if rc == 0:
            return
        if rc == _lib.ZK_ERR_INVALID_POINT:
            raise InvalidPoint(bad_index)
        raise ZkError(rc, f"{self._lib.zk_status_str(rc).decode()}: {self._lib.zk_mgpu_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.zk_mgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gather(self, mode: str):
        self._check(self._lib.zk_mgpu_set_gather(self._h, {"peer": 0, "nccl": 1}[mode]))

    def set_priority(self, high: bool):
        """Recreate the context's streams with the highest (or default) stream priority; call while idle."""
        self._check(self._lib.zk_ctx_set_priority(self._h, 1 if high else 0))

    def sum_compressed(self, points) -> Optional["CompressedRistretto"]:
        """Sum of up to 1024 compressed points in one small launch (`iter.sum()` over decompressed points, compressed);
        None if an encoding is invalid."""
        h, g = _join32(points)
        out = C.create_string_buffer(32)
        rc = self._lib.zk_sum_compressed(self._h, _ptr(h), g, out)
        if rc == _lib.ZK_ERR_INVALID_POINT:
            return None
        self._check(rc)
        return CompressedRistretto(out.raw)

    def set_wait(self, mode: int):
        """0 = spin while waiting for the device (default), 1 = block (no CPU while waiting, ~0.3 ms more latency)."""
        self._check(self._lib.zk_ctx_set_wait(self._h, mode))

    def set_staging(self, mode: int):
        self._check(self._lib.zk_mgpu_set_staging(self._h, mode))

    def set_wait(self, mode: int):
        self._check(self._lib.zk_mgpu_set_wait(self._h, mode))

    @property
    def launch_count(self) -> int:
        return int(self._lib.zk_mgpu_launch_count(self._h))

    def optional_multiscalar_mul(self, scalars, points) -> Optional[CompressedRistretto]:
        """sum scalars[i] * decompress(points[i]) over all devices; None if any encoding is invalid."""
        hs, ns = _join32(scalars)
        hp, np_ = _join32(points)
        if ns != np_:
            raise ValueError("scalars/points length mismatch")
        out = C.create_string_buffer(32)
        rc = self._lib.zk_mgpu_msm_vartime(self._h, _ptr(hs), _ptr(hp), ns, out)
        if rc == _lib.ZK_ERR_INVALID_POINT:
            return None
        self._check(rc)
        return CompressedRistretto(out.raw)

    def mixed_multiscalar_mul(self, static_scalars, table: Optional[MultiGpuTable], dyn_scalars, dyn_points,
                              offset: int = 0) -> Optional[CompressedRistretto]:
        """Cached (sharded) generators first, then the proof's own compressed points; None if a dynamic encoding is invalid."""
        hs, ns = _join32(static_scalars)
        hd, nd = _join32(dyn_scalars)
        hp, np_ = _join32(dyn_points)
        if nd != np_:
            raise ValueError("dynamic scalars/points length mismatch")
        out = C.create_string_buffer(32)
        rc = self._lib.zk_mgpu_msm_vartime_mixed(self._h, _ptr(hs), table._h if table is not None else None, offset, ns,
                                                 _ptr(hd), _ptr(hp), nd, out)
        if rc == _lib.ZK_ERR_INVALID_POINT:
            return None
        self._check(rc)
        return CompressedRistretto(out.raw)

    def vartime_multiscalar_mul(self, scalars, table: MultiGpuTable, offset: int = 0, n: Optional[int] = None) -> CompressedRistretto:
        hs, ns = _join32(scalars)
        n = ns if n is None else n
        if ns != n:
            raise ValueError("scalars/points length mismatch")
        out = C.create_string_buffer(32)
        self._check(self._lib.zk_mgpu_msm_vartime_table(self._h, _ptr(hs), table._h, offset, n, out))
        return CompressedRistretto(out.raw)
