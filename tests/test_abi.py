"""The C-ABI library loads and exports every symbol include/zkmsm.h declares; without a GPU every entry
point that needs one fails loudly (no CPU fallback).  CPU only -- no compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "zkmsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from zkvm_b200 import _lib
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zkmsm.h but not exported by libzkmsm.so"
    assert sorted(s[0] for s in _lib.SYMBOLS) == names, "ctypes table and header disagree"


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "zkmsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)      # signatures only, not the prose
    assert "torch" not in src.lower() and "at::" not in src and "#include <cuda" not in src


def test_host_only_entry_points():
    from zkvm_b200 import _lib
    lib = _lib.load()
    assert lib.zk_abi_version() == 2
    assert lib.zk_encoding_is_identity(bytes(32)) == 1
    assert lib.zk_encoding_is_identity(bytes([1]) + bytes(31)) == 0
    assert b"invalid" in lib.zk_status_str(_lib.ZK_ERR_INVALID_POINT)
    for n in (1, 1 << 10, 1 << 16, 1 << 20, 1 << 24):
        assert 4 <= lib.zk_pick_window(n) <= 20


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "zkvm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "msm_oracle" not in txt and "libsodium" not in txt.lower(), f


def test_context_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is for GPU-less hosts")
    import zkvm_b200 as zk
    with pytest.raises(zk.ZkError) as e:
        zk.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_rust_sys_matches_header():
    """The pre-staged (uncompiled: no rustc in this image) zkmsm-sys crate declares exactly the header's functions, with
    the same arity and the same pointer/integer class per parameter as the ctypes table the built library is bound with."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_rust_sys as g
    from zkvm_b200 import _lib
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]).returncode == 0
    funcs = g.parse_header()
    assert sorted(f[0] for f in funcs) == header_functions()
    rs = open(os.path.join(ROOT, "rust", "zkmsm-sys", "src", "lib.rs")).read()
    decl = dict((m.group(1), m.group(2)) for m in re.finditer(r"pub fn (zk_[a-z0-9_]+)\(([^)]*)\)", rs))
    assert sorted(decl) == header_functions()
    table = {n: (res, args) for n, res, args in _lib.SYMBOLS}
    for name, ret, params in funcs:
        rust_params = [p for p in decl[name].split(",") if p.strip()]
        assert len(rust_params) == len(params) == len(table[name][1]), name
        for (ctype, _), rp, ct in zip(params, rust_params, table[name][1]):
            is_ptr_c = "*" in ctype
            is_ptr_rs = "*const" in rp or "*mut" in rp
            is_ptr_ct = ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "_type_") and isinstance(ct._type_, type)
            assert is_ptr_c == is_ptr_rs == bool(is_ptr_ct), (name, ctype, rp, ct)
            if not is_ptr_c:
                size = {"c_int": 4, "usize": 8, "u64": 8, "f32": 4, "f64": 8}[rp.split(":")[1].strip()]
                assert C.sizeof(ct) == size, (name, ctype, rp)


def test_library_sets_hardware_queue_default_only_when_unset():
    """Loading libzkmsm.so gives CUDA_DEVICE_MAX_CONNECTIONS a default of 32 (four streams per context, several contexts
    in flight: DESIGN.md section 5.4) and never overrides the application's own setting."""
    import subprocess
    import sys
    from zkvm_b200 import _lib
    code = ("import ctypes, sys; ctypes.CDLL(sys.argv[1]); libc = ctypes.CDLL(None); libc.getenv.restype = ctypes.c_char_p; "
            "print((libc.getenv(b'CUDA_DEVICE_MAX_CONNECTIONS') or b'').decode())")
    env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}
    out = subprocess.run([sys.executable, "-c", code, _lib.LIB_PATH], env=env, capture_output=True, text=True, check=True)
    assert out.stdout.strip() == "32"
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "8"
    out = subprocess.run([sys.executable, "-c", code, _lib.LIB_PATH], env=env, capture_output=True, text=True, check=True)
    assert out.stdout.strip() == "8"
