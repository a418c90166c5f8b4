"""ctypes binding of oracle/kemr_oracle.c (TEST INFRASTRUCTURE ONLY, see oracle/oracle.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkemr_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_SO)
    return _lib


def _bits(x):
    x = np.asarray(x)
    if x.dtype == np.uint16:
        return np.ascontiguousarray(x)
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = (x.view(np.uint32) >> np.uint32(16)).astype(np.uint16)
    assert np.array_equal((b.astype(np.uint32) << np.uint32(16)).view(np.float32), x), "values must be bf16-representable"
    return b


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _csr(hits):
    if hits is None:
        return None, None, None
    rp, col, bon = hits
    return (np.ascontiguousarray(rp, np.int64), np.ascontiguousarray(col, np.int32),
            np.ascontiguousarray(bon, np.float64))


class _QueryWeights:
    """Per-query fusion weights (arrays [Q]) for the duration of one oracle call; scalars pass through."""

    def __init__(self, w_a, w_b, Q):
        self.arrays = None
        if np.ndim(w_a) or np.ndim(w_b):
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(w_a, np.float64).reshape(-1), (Q,)))
            b = np.ascontiguousarray(np.broadcast_to(np.asarray(w_b, np.float64).reshape(-1), (Q,)))
            self.arrays = (a, b)
        self.w_a = 0.0 if self.arrays else float(w_a)
        self.w_b = 0.0 if self.arrays else float(w_b)

    def __enter__(self):
        if self.arrays:
            load().oracle_set_query_weights(_p(self.arrays[0]), _p(self.arrays[1]))
        return self

    def __exit__(self, *exc):
        if self.arrays:
            load().oracle_set_query_weights(None, None)


def scores(q, ga, gb=None, w_a=1.0, w_b=0.0, alpha=1.0, hits=None):
    q, ga = _bits(q), _bits(ga)
    gb = _bits(gb) if gb is not None else None
    Q, D = q.shape
    M = ga.shape[0]
    out = np.empty((Q, M), np.float64)
    rp, col, bon = _csr(hits)
    with _QueryWeights(w_a, w_b, Q) as w:
        load().oracle_scores(_p(q), Q, _p(ga), _p(gb), C.c_int64(M), D, C.c_double(w.w_a), C.c_double(w.w_b),
                             C.c_double(alpha), _p(rp), _p(col), _p(bon), _p(out))
    return out


def topk_rank(q, ga, gb=None, w_a=1.0, w_b=0.0, alpha=1.0, hits=None, k=0, target=None):
    """(idx int64 [Q,k], score f64 [Q,k], rank int64 [Q]) -- idx/score None if k == 0, rank None if no target."""
    q, ga = _bits(q), _bits(ga)
    gb = _bits(gb) if gb is not None else None
    Q, D = q.shape
    M = ga.shape[0]
    idx = np.empty((Q, k), np.int64) if k else None
    sc = np.empty((Q, k), np.float64) if k else None
    tgt = np.ascontiguousarray(target, np.int64) if target is not None else None
    rank = np.empty((Q,), np.int64) if target is not None else None
    rp, col, bon = _csr(hits)
    with _QueryWeights(w_a, w_b, Q) as w:
        load().oracle_topk_rank(_p(q), Q, _p(ga), _p(gb), C.c_int64(M), D, C.c_double(w.w_a), C.c_double(w.w_b),
                                C.c_double(alpha), _p(rp), _p(col), _p(bon), k, _p(idx), _p(sc), _p(tgt), _p(rank))
    return idx, sc, rank
