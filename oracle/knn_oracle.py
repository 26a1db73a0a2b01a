"""ctypes wrapper of oracle/knn_oracle.c.  TEST INFRASTRUCTURE ONLY (see the C file's header).

Mirrors the reference's call protocol (main.py:268-269, :200, :300):
``fit(E)`` then ``kneighbors(q, n_neighbors=k) -> (dist f32 [nq,k], ind i64 [nq,k])``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "_build", "libknn_oracle.so")
    src = os.path.join(_HERE, "knn_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        L.knn_normalize_rows.argtypes = [f32p, f32p, ctypes.c_int64, ctypes.c_int]
        L.knn_normalize_rows.restype = None
        L.knn_cosine_topk.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int64, f32p, i64p]
        L.knn_cosine_topk.restype = None
        L.knn_merge_topk.argtypes = [f32p, i64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, i64p]
        L.knn_merge_topk.restype = None
        _LIB = L
    return _LIB


def _f32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i64(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def normalize_rows(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    _lib().knn_normalize_rows(_f32(x), _f32(out), x.shape[0], x.shape[1])
    return out


def cosine_topk(ehat: np.ndarray, qhat: np.ndarray, k: int, idx_base: int = 0):
    ehat = np.ascontiguousarray(ehat, dtype=np.float32)
    qhat = np.ascontiguousarray(qhat, dtype=np.float32)
    nq = qhat.shape[0]
    dist = np.empty((nq, k), dtype=np.float32)
    idx = np.empty((nq, k), dtype=np.int64)
    _lib().knn_cosine_topk(_f32(ehat), ehat.shape[0], ehat.shape[1], _f32(qhat), nq, k, idx_base,
                           _f32(dist), _i64(idx))
    return dist, idx


def merge_topk(dist_parts: np.ndarray, idx_parts: np.ndarray):
    dist_parts = np.ascontiguousarray(dist_parts, dtype=np.float32)
    idx_parts = np.ascontiguousarray(idx_parts, dtype=np.int64)
    n_parts, nq, k = dist_parts.shape
    dist = np.empty((nq, k), dtype=np.float32)
    idx = np.empty((nq, k), dtype=np.int64)
    _lib().knn_merge_topk(_f32(dist_parts), _i64(idx_parts), n_parts, nq, k, _f32(dist), _i64(idx))
    return dist, idx


class OracleNearestNeighbors:
    """sklearn-protocol object over the C oracle (the contract order: dist asc, index asc)."""

    def __init__(self, n_neighbors: int = 16, metric: str = "cosine", algorithm: str = "brute"):
        if metric != "cosine":
            raise ValueError("only metric='cosine' is on the reference's path (main.py:268)")
        self.n_neighbors = n_neighbors

    def fit(self, X):
        self._ehat = normalize_rows(np.asarray(X))
        return self

    def kneighbors(self, X, n_neighbors=None, return_distance=True):
        k = self.n_neighbors if n_neighbors is None else n_neighbors
        if k > self._ehat.shape[0]:
            raise ValueError("Expected n_neighbors <= n_samples_fit")   # sklearn's error, _base.py
        q = normalize_rows(np.asarray(X).reshape(-1, self._ehat.shape[1]))
        dist, idx = cosine_topk(self._ehat, q, k)
        return (dist, idx) if return_distance else idx
