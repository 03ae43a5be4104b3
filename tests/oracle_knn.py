"""ctypes access to the C k-NN oracle (oracle/knn_oracle.c).  Test infrastructure only."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libknn_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            import subprocess

            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
        _lib = ctypes.CDLL(_SO)
        _lib.knn_canonical_dot.restype = ctypes.c_float
    return _lib


def _vp(x):
    return x.ctypes.data_as(ctypes.c_void_p)


def canonical_dot(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return lib().knn_canonical_dot(_vp(a), _vp(b), a.shape[0])


def topk(gallery: np.ndarray, queries: np.ndarray, k: int, index_base: int = 0, threads: int = 8,
         return_all: bool = False):
    gallery = np.ascontiguousarray(gallery, np.float32)
    queries = np.ascontiguousarray(queries, np.float32)
    n, d = gallery.shape
    q = queries.shape[0]
    idx = np.zeros((q, k), np.int64)
    score = np.zeros((q, k), np.float32)
    allsc = np.zeros((q, n), np.float32) if return_all else None
    lib().knn_oracle_topk(_vp(gallery), n, d, _vp(queries), q, k, ctypes.c_int64(index_base), _vp(idx), _vp(score),
                          _vp(allsc) if return_all else None, threads)
    return (idx, score, allsc) if return_all else (idx, score)
