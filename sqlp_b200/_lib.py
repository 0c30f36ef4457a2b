"""ctypes binding of ``libsqlp_b200.so`` (the C ABI in ``include/sqlp_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``sqlp_b200.build()``.
There is no CPU fallback: if the shared library is missing, or no CUDA device is present,
the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SO_PATH = os.environ.get("SQLP_B200_LIB") or os.path.join(_HERE, "libsqlp_b200.so")   # override: kernel-variant sweeps
SRC = os.path.join(_HERE, "csrc", "sqlp_api.cu")
HEADER = os.path.join(ROOT, "include", "sqlp_b200.h")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

OK, E_INVALID, E_CUDA, E_UNSUPPORTED, E_NO_ARGMAX, E_NOMEM, E_NCCL, E_RANGE, E_IO = 0, -1, -2, -3, -4, -5, -6, -7, -8
MIN_SENSE, MAX_SENSE = 0, 1


class SqlpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[sqlp {code}] {msg}")
        self.code = code


class NoArgmaxError(SqlpError):
    """The reference throws UndefRefError here (epigraph.jl:140)."""


def sources():
    d = os.path.join(_HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [HEADER]


def source_id() -> str:
    """Content hash of every source of the library and of the compile flags: the binary carries it
    (``sqlp_version()``), so a stale .so is recognised whatever its timestamps say."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sources():
        with open(p, "rb") as fh:
            h.update(os.path.basename(p).encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def built_id(path: str = None) -> str | None:
    """The source id compiled into an existing binary (read from the file, no CUDA needed)."""
    path = path or SO_PATH
    if not os.path.exists(path):
        return None
    with open(path, "rb") as fh:
        blob = fh.read()
    i = blob.find(b"sqlp-build-id:")
    return blob[i + 14:i + 30].decode("ascii", "replace") if i >= 0 else None


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> sqlp_b200/libsqlp_b200.so"""
    sid = source_id()
    if not force and built_id() == sid:
        return SO_PATH
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc, *NVCC_FLAGS, f"-DSQLP_BUILD_ID=\"{sid}\"", "-o", SO_PATH, SRC, "-ldl"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{r.stdout}\n{r.stderr}")
    return SO_PATH


_vp, _i32, _i64, _u64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
_P = C.POINTER

# name -> argtypes; every function returns int32 unless listed in _RESTYPE
SIGNATURES = {
    "sqlp_guard_check": [_P(_i64), _P(_i64)],
    "sqlp_ctx_create": [_i32, _P(_vp)],
    "sqlp_nccl_unique_id": [_vp],
    "sqlp_ctx_create_dist": [_i32, _i32, _i32, _vp, _P(_vp)],
    "sqlp_ctx_create_multi": [_i32, _vp, _P(_vp)],
    "sqlp_ctx_destroy": [_vp],
    "sqlp_ctx_set_stream": [_vp, _vp],
    "sqlp_ctx_synchronize": [_vp],
    "sqlp_ctx_launch_count": [_vp, _P(_i64)],
    "sqlp_ctx_timer_start": [_vp],
    "sqlp_ctx_timer_stop": [_vp],
    "sqlp_ctx_timer_elapsed_ms": [_vp, _P(_f64)],
    "sqlp_ctx_profile": [_vp, _i32],
    "sqlp_ctx_profile_read": [_vp, _i32, _P(_f64), _P(_i64), _P(_f64)],
    "sqlp_ctx_profile_classes": [_vp, _i32, _vp, _vp, _vp],
    "sqlp_ctx_set_screen": [_vp, _i32],
    "sqlp_epi_screen_stats": [_vp, _vp],
    "sqlp_cell_build_cuts2_dev": [_i32, _vp, _vp, _vp],
    "sqlp_epi_view_columns": [_vp, _P(_i64), _P(_i64)],
    "sqlp_pool_create": [_vp, _i64, _P(_vp)],
    "sqlp_pool_destroy": [_vp],
    "sqlp_pool_push": [_vp, _vp, _P(_i32), _P(_i64)],
    "sqlp_pool_push_batch": [_vp, _i64, _vp, _vp, _vp],
    "sqlp_pool_push_dev": [_vp, _i64, _vp],
    "sqlp_pool_size": [_vp, _P(_i64)],
    "sqlp_pool_get": [_vp, _i64, _vp],
    "sqlp_pool_hash": [_vp, _vp, _P(_u64)],
    "sqlp_epi_create": [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _P(_vp)],
    "sqlp_epi_destroy": [_vp],
    "sqlp_epi_add_scenarios": [_vp, _i64, _vp, _vp],
    "sqlp_epi_add_scenarios_dev": [_vp, _i64, _vp, _vp],
    "sqlp_epi_set_outcomes": [_vp, _i64, _vp, _vp, _vp],
    "sqlp_epi_set_distributions": [_vp, _vp, _vp, _vp],
    "sqlp_epi_sample_scenarios": [_vp, _i64, _u64, _u64],
    "sqlp_epi_counts": [_vp, _P(_i64), _P(_i64), _P(_f64)],
    "sqlp_epi_delta": [_vp, _i64, _vp, _vp],
    "sqlp_epi_argmax": [_vp, _vp, _i32, _vp, _vp],
    "sqlp_epi_build_cut": [_vp, _vp, _P(_f64), _vp, _P(_f64), _P(_f64)],
    "sqlp_epi_build_cuts2": [_vp, _vp, _vp, _vp, _vp, _P(_f64), _vp],
    "sqlp_cell_build_cuts2": [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "sqlp_epi_build_cuts2_dev": [_vp, _vp, _vp],
    "sqlp_cell_sd_step": [_i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "sqlp_eval_dual": [_vp, _i64, _i64, _vp, _P(_f64)],
    "sqlp_epi_set_weights": [_vp, _f64, _f64],
    "sqlp_epi_cuts_push": [_vp, _f64, _vp, _f64],
    "sqlp_epi_cuts_set_incumbent": [_vp, _f64, _vp, _f64],
    "sqlp_epi_cuts_commit": [_vp, _i32],
    "sqlp_epi_cuts_delete": [_vp, _i64, _vp],
    "sqlp_epi_cuts_count": [_vp, _P(_i64), _P(_i32)],
    "sqlp_epi_cuts_get": [_vp, _i64, _P(_f64), _vp, _P(_f64)],
    "sqlp_epi_evaluate": [_vp, _vp, _i32, _P(_f64)],
    "sqlp_epi_master_rows": [_vp, _vp, _P(_i64)],
    "sqlp_cell_check_improvement": [_i32, _vp, _vp, _vp, _vp, _f64, _vp],
    "sqlp_smps_load": [C.c_char_p, C.c_char_p, C.c_char_p, _P(_vp)],
    "sqlp_smps_destroy": [_vp],
    "sqlp_smps_dims": [_vp, _vp],
    "sqlp_smps_name": [_vp, _i32, _i64, C.c_char_p, _i64],
    "sqlp_smps_cor": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "sqlp_smps_stage2": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "sqlp_smps_elements": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "sqlp_epi_create_smps": [_vp, _vp, _vp, _P(_vp)],
}
_RESTYPE = {"sqlp_version": C.c_char_p, "sqlp_last_error": C.c_char_p}

_lib = None


def lib():
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise SqlpError(E_CUDA, f"{SO_PATH} is missing: run `python -c 'import "
                                    "__graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        if not os.environ.get("SQLP_B200_LIB") and built_id() != source_id():
            raise SqlpError(E_CUDA, f"{SO_PATH} was built from other sources (binary {built_id()}, tree "
                                    f"{source_id()}): rebuild with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(SO_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _i32
        for name, rt in _RESTYPE.items():
            fn = getattr(L, name)
            fn.argtypes = []
            fn.restype = rt
        _lib = L
    return _lib


def check(status: int):
    if status == OK:
        return
    msg = lib().sqlp_last_error().decode("utf-8", "replace")
    if status == E_NO_ARGMAX:
        raise NoArgmaxError(status, msg)
    raise SqlpError(status, msg)


def declared_symbols():
    """Every SQLP_API function name declared in include/sqlp_b200.h."""
    import re
    with open(HEADER) as fh:
        text = fh.read()
    return sorted(set(re.findall(r"SQLP_API\s+[\w\s\*]+?\b(sqlp_\w+)\s*\(", text)))
