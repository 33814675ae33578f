"""ctypes loader for oracle/_build/liboracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "tfbs_oracle.c")
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        L.oracle_encode.argtypes = [vp, i64, vp, vp]
        L.oracle_encode.restype = None
        L.oracle_base_counts.argtypes = [vp, i64, vp]
        L.oracle_base_counts.restype = None
        L.oracle_pwm_calculate.argtypes = [vp, i64, vp, i32, vp]
        L.oracle_pwm_calculate.restype = None
        L.oracle_pwm_calculate_mt.argtypes = [vp, i64, vp, i32, vp, i32]
        L.oracle_pwm_calculate_mt.restype = None
        L.oracle_pwm_revcomp.argtypes = [vp, i32, vp]
        L.oracle_pwm_revcomp.restype = None
        L.oracle_pwm_search.argtypes = [vp, i64, vp, vp, i32, i32, vp, vp, vp, vp, vp, i64]
        L.oracle_pwm_search.restype = i64
        L.oracle_pwm_search_mt.argtypes = [vp, i64, vp, vp, i32, i32, vp, vp, vp, vp, vp, i64, i32]
        L.oracle_pwm_search_mt.restype = i64
        L.oracle_pwm_scan_count.argtypes = [vp, i64, vp, vp, i32, i32, vp, i32, vp, vp]
        L.oracle_pwm_scan_count.restype = i64
        L.oracle_shape_tracks.argtypes = [vp, i64, vp, vp, i32, vp]
        L.oracle_shape_tracks.restype = None
        L.oracle_shape_tracks_set.argtypes = [vp, vp, vp, i64, i64, vp, vp, i32, i32, vp]
        L.oracle_shape_tracks_set.restype = None
        _lib = L
    return _lib


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def as_bytes(seq) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    a = np.ascontiguousarray(seq, dtype=np.uint8)
    return a
