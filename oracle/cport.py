"""ctypes loader/builder for oracle/c/ame_oracle.c (TEST INFRASTRUCTURE - see oracle/README.md)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "c", "ame_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libame_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc the C twin into oracle/_build/ (git-ignored, travels to the GPU box)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            try:
                build()
            except Exception:
                return None
        lib = ctypes.CDLL(_SO)
        i16p = ctypes.POINTER(ctypes.c_int16)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.ame_oracle_compress.argtypes = [i16p, i16p, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_double, ctypes.c_double, ctypes.c_double, dp]
        lib.ame_oracle_compress.restype = ctypes.c_int
        lib.ame_oracle_window_rms.argtypes = [i16p, ctypes.POINTER(ctypes.c_uint16), ctypes.c_int64, ctypes.c_int64]
        lib.ame_oracle_window_rms.restype = ctypes.c_int
        lib.ame_oracle_kfilter_df2.argtypes = [i16p, dp, ctypes.c_int64, dp, dp]
        lib.ame_oracle_kfilter_df2.restype = ctypes.c_int
        lib.ame_oracle_alimiter.argtypes = [i16p, i16p, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_double, dp]
        lib.ame_oracle_alimiter.restype = ctypes.c_int
        _lib = lib
    return _lib


def available() -> bool:
    return _load() is not None


def compress(pcm, fs, threshold, ratio, attack=5.0, release=50.0, return_att=False):
    lib = _load()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty_like(pcm)
    n = pcm.shape[0]
    att = np.empty(n, dtype=np.float64) if return_att else None
    lib.ame_oracle_compress(pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)),
                            out.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), n, float(fs),
                            float(threshold), float(ratio), float(attack), float(release),
                            att.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if return_att else None)
    return (out, att) if return_att else out


def window_rms(pcm, look):
    lib = _load()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty(pcm.shape[0], dtype=np.uint16)
    lib.ame_oracle_window_rms(pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)),
                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), pcm.shape[0], int(look))
    return out


def kfilter_df2(pcm, b, a):
    lib = _load()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty(pcm.shape, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ame_oracle_kfilter_df2(pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), out.ctypes.data_as(dp),
                               pcm.shape[0], b.ctypes.data_as(dp), a.ctypes.data_as(dp))
    return out


def alimiter(pcm, fs, limit=0.98, attack_ms=5.0, release_ms=50.0, return_att=False):
    lib = _load()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty_like(pcm)
    n = pcm.shape[0]
    att = np.empty(n, dtype=np.float64) if return_att else None
    rc = lib.ame_oracle_alimiter(pcm.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)),
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), n, float(fs), float(limit),
                                 float(attack_ms), float(release_ms),
                                 att.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if return_att else None)
    if rc:
        raise ValueError("alimiter: attack too short for this sample rate")
    return (out, att) if return_att else out
