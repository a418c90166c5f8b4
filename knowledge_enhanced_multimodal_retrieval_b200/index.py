"""Device-resident gallery ("index") and its search call -- the B200 replacement for the scan
the reference delegates to remote code (`CLIPRetriever.search`, fetched from the HF hub at
`src/clip/clip_retrieval.py:15-23`).  Two flavours:

* `GalleryIndex`  -- galleries live in torch CUDA tensors; queries may already be on the device.
* `HostIndex`     -- thin wrapper over the C handle `kemr_index_*`: HOST buffers in, HOST buffers
                     out, copies and synchronisation inside the call (the end-to-end path).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine
from ._lib import KemrError


def _bits(x) -> np.ndarray:
    """fp32 host array of bf16-representable values, or bf16 bit patterns -> contiguous uint16 bits."""
    from .synth import f32_to_bf16_bits
    x = np.asarray(x)
    if x.dtype == np.uint16:
        return np.ascontiguousarray(x)
    return f32_to_bf16_bits(np.ascontiguousarray(x, dtype=np.float32))


class GalleryIndex:
    """Row-major bf16 galleries resident in HBM: image (T2I) and optional target-text (T2T)."""

    def __init__(self, image_embeddings, target_embeddings=None, uuids: Optional[Sequence[str]] = None,
                 idx_base: int = 0):
        self.image = engine.quantize(image_embeddings)
        self.target = engine.quantize(target_embeddings) if target_embeddings is not None else None
        if self.target is not None and self.target.shape != self.image.shape:
            raise KemrError("image and target galleries must have the same shape")
        self.uuids = list(uuids) if uuids is not None else None
        if self.uuids is not None and len(self.uuids) != self.image.shape[0]:
            raise KemrError("uuids length != gallery rows")
        self.idx_base = idx_base
        self._uuid_to_row: Optional[Dict[str, int]] = None

    @property
    def M(self) -> int:
        return self.image.shape[0]

    @property
    def D(self) -> int:
        return self.image.shape[1]

    def row_of(self, uuid: str) -> Optional[int]:
        if self._uuid_to_row is None:
            self._uuid_to_row = {u: j for j, u in enumerate(self.uuids or [])}
        return self._uuid_to_row.get(engine.uri_tail(uuid))

    def hits_from_uuid_lists(self, lists: Sequence[Sequence[str]], bonus: float) -> engine.KGHits:
        """KG result lists (uuids or URIs) per query -> CSR with `bonus` per unique known row."""
        cols = []
        for lst in lists:
            seen: Dict[int, None] = {}
            for u in lst:
                r = self.row_of(u)
                if r is not None:
                    seen[r] = None
            cols.append(list(seen))
        return engine.KGHits.from_lists(cols, [[bonus] * len(c) for c in cols])

    def search(self, query_embeddings, k: int = 10, t2i_weight: float = 1.0, t2t_weight: float = 0.0,
               alpha: float = 1.0, hits: Optional[engine.KGHits] = None, path: int = _lib.PATH_AUTO,
               normalize: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top-k rows by alpha*(w_i*T2I + w_t*T2T) + KG bonus.  Returns (idx int64, score f64) on device."""
        q = engine.quantize(query_embeddings, normalize=normalize)
        two = self.target is not None and t2t_weight != 0.0
        return engine.scan_topk(q, self.image, self.target if two else None, t2i_weight,
                                t2t_weight if two else 0.0, alpha, hits, k, idx_base=self.idx_base, path=path)


class HostIndex:
    """`kemr_index_*` handle: galleries uploaded once, `search` takes and returns host arrays."""

    def __init__(self, image_embeddings, target_embeddings=None, max_queries: int = 1024, max_k: int = 100):
        if not torch.cuda.is_available():
            raise KemrError("no CUDA device: this engine has no CPU path")
        lib = _lib.load()
        a = _bits(image_embeddings)
        b = _bits(target_embeddings) if target_embeddings is not None else None
        self.M, self.D = a.shape
        self.max_queries, self.max_k = max_queries, max_k
        h = C.c_void_p()
        _lib.check(lib.kemr_index_create(a.ctypes.data_as(C.c_void_p),
                                         b.ctypes.data_as(C.c_void_p) if b is not None else None,
                                         self.M, self.D, max_queries, max_k, C.byref(h)))
        self._h = h
        self._lib = lib
        self._source = None
        self._lanes = []

    def close(self):
        for lane in getattr(self, "_lanes", []):
            lane.close()
        self._lanes = []
        if getattr(self, "_h", None):
            self._lib.kemr_index_destroy(self._h)
            self._h = None

    def lane(self) -> "HostIndex":
        """A second lane over the same resident galleries (`kemr_index_share`): own stream, workspace and staging
        buffers.  `submit` on one lane while the other is still scanning and the next batch's queries cross PCIe
        meanwhile; closed together with this index."""
        if getattr(self, "_source", None) is not None:
            raise KemrError("lanes are created from the index that owns the galleries")
        h = C.c_void_p()
        _lib.check(self._lib.kemr_index_share(self._h, C.byref(h)))
        other = HostIndex.__new__(HostIndex)
        other._lib, other._h, other._source, other._lanes = self._lib, h, self, []
        other.M, other.D, other.max_queries, other.max_k = self.M, self.D, self.max_queries, self.max_k
        self._lanes = getattr(self, "_lanes", []) + [other]
        return other

    def submit(self, queries_f32: np.ndarray, k: int = 10, t2i_weight: float = 1.0, t2t_weight: float = 0.0,
               alpha: float = 1.0, hits_csr=None, normalize: bool = False, out=None):
        """`search` without the wait (`kemr_index_submit_host`): returns the result arrays at once; they are valid after
        `wait()`.  One submitted search per lane; page-locked buffers must stay untouched until then."""
        q = _as(queries_f32, np.float32)
        Q = q.shape[0]
        if out is None:
            out = (np.empty((Q, k), np.int64), np.empty((Q, k), np.float64), np.empty((Q,), np.int32))
        idx, score, flags = out
        rp = cc = bb = None
        if hits_csr is not None:
            rp, cc, bb = _as(hits_csr[0], np.int64), _as(hits_csr[1], np.int32), _as(hits_csr[2], np.float64)
        _lib.check(self._lib.kemr_index_submit_host(self._h, _addr(q), Q, int(normalize), float(t2i_weight),
                                                    float(t2t_weight), float(alpha), _addr(rp), _addr(cc), _addr(bb), k,
                                                    _addr(idx), _addr(score), _addr(flags)))
        self._keep = (q, rp, cc, bb, out)                 # the C side reads / writes them until wait()
        return idx, score, flags

    def wait(self):
        _lib.check(self._lib.kemr_index_wait(self._h))
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, queries_f32: np.ndarray, k: int = 10, t2i_weight: float = 1.0, t2t_weight: float = 0.0,
               alpha: float = 1.0, hits_csr=None, normalize: bool = False, out=None):
        """queries fp32 [Q, D] (host) -> (idx int64 [Q,k], score float64 [Q,k], flags int32 [Q]) host arrays.
        hits_csr = (rowptr int64, col int32, bonus float64) host arrays or None.  A uint16 array is taken as bf16 bit
        patterns (kemr_index_search_host_bf16: half the PCIe bytes, no quantise kernel; `normalize` must be False)."""
        bf16 = isinstance(queries_f32, np.ndarray) and queries_f32.dtype == np.uint16
        if bf16 and normalize:
            raise KemrError("bf16 queries cannot be re-normalised")
        q = _as(queries_f32, np.uint16 if bf16 else np.float32)
        Q = q.shape[0]
        if out is None:
            out = (np.empty((Q, k), np.int64), np.empty((Q, k), np.float64), np.empty((Q,), np.int32))
        idx, score, flags = out
        rp = cc = bb = None
        if hits_csr is not None:
            rp, cc, bb = _as(hits_csr[0], np.int64), _as(hits_csr[1], np.int32), _as(hits_csr[2], np.float64)
        if bf16:
            _lib.check(self._lib.kemr_index_search_host_bf16(self._h, _addr(q), Q, float(t2i_weight), float(t2t_weight),
                                                             float(alpha), _addr(rp), _addr(cc), _addr(bb), k,
                                                             _addr(idx), _addr(score), _addr(flags)))
        else:
            _lib.check(self._lib.kemr_index_search_host(self._h, _addr(q), Q, int(normalize), float(t2i_weight),
                                                        float(t2t_weight), float(alpha), _addr(rp), _addr(cc), _addr(bb), k,
                                                        _addr(idx), _addr(score), _addr(flags)))
        return idx, score, flags


def _as(x, dtype):
    """x as a C-contiguous array of `dtype` (no copy when it already is: the serving path calls this per request)."""
    if type(x) is np.ndarray and x.dtype == dtype and x.flags.c_contiguous:
        return x
    return np.ascontiguousarray(x, dtype=dtype)


def _addr(x):
    """Address of a numpy array's buffer for a `void*` argument (`arr.ctypes.data_as` builds two Python objects per
    call -- 17 us of an 83 us batch-1 request went to seven of them)."""
    return None if x is None else x.__array_interface__["data"][0]
