"""ctypes wrapper of oracle/mmr_oracle.c.  TEST INFRASTRUCTURE ONLY (see the C file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libmmr_oracle.so")
        src = os.path.join(_HERE, "mmr_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = ctypes.CDLL(so)
        L.mmr_rerank.restype = ctypes.c_int
        L.mmr_rerank.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                 ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
        _LIB = L
    return _LIB


def mmr_rerank(emb: np.ndarray, scores: np.ndarray, emb_idx: np.ndarray, lambda_param: float, top_k: int = 20) -> np.ndarray:
    """Positions (into the ranked candidate list) chosen by the reference's rerank_with_mmr (main.py:133-169)."""
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    emb_idx = np.ascontiguousarray(emb_idx, dtype=np.int64)
    out = np.full(max(top_k, 1), -1, dtype=np.int32)
    n = _lib().mmr_rerank(emb.ctypes.data, emb.shape[0], emb.shape[1], scores.ctypes.data, emb_idx.ctypes.data,
                          scores.shape[0], float(lambda_param), int(top_k), out.ctypes.data)
    return out[:n]
