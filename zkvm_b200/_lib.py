"""ctypes binding of libzkmsm.so (include/zkmsm.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzkmsm.so")
# Developer-only A/B hook (tools/ab_build.sh): a variant build is honoured only when ZKMSM_DEV=1 is ALSO set and the
# file lives inside this package directory, so a stray environment variable can never redirect the product path.
if os.environ.get("ZKMSM_DEV") == "1" and os.environ.get("ZKMSM_LIB"):
    _cand = os.path.abspath(os.environ["ZKMSM_LIB"])
    if os.path.dirname(_cand) == _HERE and os.path.basename(_cand).startswith("libzkmsm"):
        LIB_PATH = _cand

ZK_OK, ZK_ERR_CUDA, ZK_ERR_INVALID_POINT, ZK_ERR_ARG, ZK_ERR_NOMEM = 0, -1, -2, -3, -4

# every symbol include/zkmsm.h declares: (name, restype, argtypes)
_vp, _sz, _u8p, _i = C.c_void_p, C.c_size_t, C.c_char_p, C.c_int
SYMBOLS = [
    ("zk_abi_version", _i, []),
    ("zk_status_str", C.c_char_p, [_i]),
    ("zk_last_error", C.c_char_p, [_vp]),
    ("zk_ctx_create", _i, [_i, C.POINTER(_vp)]),
    ("zk_ctx_destroy", None, [_vp]),
    ("zk_ctx_sync", _i, [_vp]),
    ("zk_ctx_set_wait", _i, [_vp, _i]),
    ("zk_ctx_stream", _vp, [_vp]),
    ("zk_table_create", _i, [_vp, _sz, C.POINTER(_vp)]),
    ("zk_table_destroy", None, [_vp]),
    ("zk_table_len", _sz, [_vp]),
    ("zk_table_capacity", _sz, [_vp]),
    ("zk_table_clear", None, [_vp]),
    ("zk_table_append_compressed", _i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_table_append_compressed_dev", _i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_table_append_uniform", _i, [_vp, _vp, _vp, _sz]),
    ("zk_table_append_uniform_dev", _i, [_vp, _vp, _vp, _sz]),
    ("zk_table_append_extended", _i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_table_append_extended_dev", _i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_table_precompute", _i, [_vp, _vp, _i]),
    ("zk_table_precomputed_window", _i, [_vp]),
    ("zk_table_compress", _i, [_vp, _vp, _sz, _sz, _vp]),
    ("zk_table_compress_dev", _i, [_vp, _vp, _sz, _sz, _vp]),
    ("zk_msm_vartime", _i, [_vp, _vp, _vp, _sz, _vp]),
    ("zk_msm_vartime_table", _i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    ("zk_msm_vartime_mixed", _i, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp]),
    ("zk_msm_vartime_batch", _i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    ("zk_msm_vartime_table_batch", _i, [_vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    ("zk_msm_table_dev", _i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    ("zk_ext_sum_compress_dev", _i, [_vp, _vp, _sz, _vp]),
    ("zk_sum_compressed", _i, [_vp, _vp, _sz, _vp]),
    ("zk_ctx_set_priority", _i, [_vp, _i]),
    ("zk_encoding_is_identity", _i, [_vp]),
    ("zk_ctx_set_window", _i, [_vp, _i]),
    ("zk_pick_window", _i, [_sz]),
    ("zk_bench_int_pipe", _i, [_vp, _i, C.POINTER(C.c_double)]),
    ("zk_ctx_set_profiling", _i, [_vp, _i]),
    ("zk_ctx_last_phase_ms", _i, [_vp, C.POINTER(C.c_float * 4)]),
    ("zk_ctx_launch_count", C.c_uint64, [_vp]),
    ("zk_host_register", _i, [_vp, _sz]),
    ("zk_host_unregister", _i, [_vp]),
    ("zk_ctx_set_staging", _i, [_vp, _i]),
    ("zk_ctx_staged_bytes", C.c_uint64, [_vp]),
    ("zk_table_append_extended_unchecked", _i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_table_append_extended_unchecked_dev", _i, [_vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_mgpu_create", _i, [C.POINTER(_i), _i, C.POINTER(_vp)]),
    ("zk_mgpu_destroy", None, [_vp]),
    ("zk_mgpu_device_count", _i, [_vp]),
    ("zk_mgpu_last_error", C.c_char_p, [_vp]),
    ("zk_mgpu_set_gather", _i, [_vp, _i]),
    ("zk_mgpu_set_staging", _i, [_vp, _i]),
    ("zk_mgpu_set_wait", _i, [_vp, _i]),
    ("zk_mgpu_launch_count", C.c_uint64, [_vp]),
    ("zk_mgpu_msm_vartime", _i, [_vp, _vp, _vp, _sz, _vp]),
    ("zk_mgpu_table_create", _i, [_vp, _sz, C.POINTER(_vp)]),
    ("zk_mgpu_table_destroy", None, [_vp]),
    ("zk_mgpu_table_len", _sz, [_vp]),
    ("zk_mgpu_table_clear", None, [_vp]),
    ("zk_mgpu_table_append_compressed", _i, [_vp, _vp, _sz, C.POINTER(_sz)]),
    ("zk_mgpu_table_append_uniform", _i, [_vp, _vp, _sz]),
    ("zk_mgpu_msm_vartime_table", _i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    ("zk_mgpu_msm_vartime_mixed", _i, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp]),
]

_lib = None


class ZkError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        super().__init__(f"zkmsm status {status}: {detail}")


def load() -> C.CDLL:
    """Load libzkmsm.so (built in-tree by `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
            "zkvm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError here == ABI drift between header and library
        fn.restype, fn.argtypes = res, args
    if lib.zk_abi_version() != 2:
        raise ImportError("libzkmsm.so ABI version mismatch")
    _lib = lib
    return lib
