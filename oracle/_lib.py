"""Build + ctypes binding of ``oracle/xrs_oracle.c`` (test infrastructure only)."""

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "xrs_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "libxrs_oracle.so")
_lock = threading.Lock()
_lib = None

c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
c_p = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (no FMA contraction, OpenMP)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (
        not force
        and os.path.exists(_OUT)
        and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC)
    ):
        return _OUT
    cmd = [
        "gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
        "-fvisibility=hidden", "-o", _OUT, _SRC, "-lm",
    ]
    subprocess.run(cmd, check=True)
    return _OUT


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            path = build()
            L = ctypes.CDLL(path)
            L.xrso_num_threads.restype = ctypes.c_int
            L.xrso_set_num_threads.argtypes = [ctypes.c_int]
            L.xrso_ij_bboxes.argtypes = [c_p, c_p, c_i64, c_i64, c_p, c_i64, c_f64, c_i64, c_p]
            L.xrso_rectify_ij_block.argtypes = [
                c_p, c_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_p, c_i64, c_i64, c_i64, c_i64,
                c_f64, c_f64, c_f64, c_f64, c_f64,
            ]
            L.xrso_rectify_ij.argtypes = [
                c_p, c_p, c_i64, c_i64, c_p, c_p, c_i64, c_i64, c_i64, c_i64,
                c_f64, c_f64, c_f64, c_f64, c_f64, ctypes.c_int, c_f64,
            ]
            L.xrso_gather_ij.argtypes = [
                c_p, ctypes.c_int, c_i64, c_i64, c_i64, c_p, c_p, c_i64, c_i64, ctypes.c_int,
            ]
            L.xrso_gather_ij.restype = ctypes.c_int
            L.xrso_mode.argtypes = [c_p, c_i64, c_i64, c_i64, c_i64, c_p]
            _lib = L
        return _lib


def ptr(a) -> int:
    return a.ctypes.data
