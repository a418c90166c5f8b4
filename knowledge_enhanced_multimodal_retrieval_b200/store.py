"""The data either side of the scan: persisted bf16 gallery shards, the uuid <-> row map, the
Text2SPARQL result files and the KG-hit CSR the scan consumes.

Reference counterparts: the `data/embeddings` directory the remote retriever loads (`clip_retrieval.py:28,35`), the
uuid lists the evaluator carries next to the embeddings (`evaluator.py:143,183-184`), one text file per query under
`experiments/text2sparql/results` with one artefact URI per line (`evaluator.py:43-50`), the URI -> uuid cut
(`fusion.py:76`, `text2sparql_retrieval.py:57`) and the per-strategy hit loops (`fusion.py:68-80,119-130,180-204`).

Layout of a store directory:
    image.kemr / target.kemr   64-byte header + M*D bf16 row-major (kemr_store_write / kemr_store_load)
    uuids.txt                  one artefact uuid per line, row order
A rank of a row-sharded gallery reads only its byte range of the two files.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, engine
from ._lib import KemrError
from .distributed import shard_bounds
from .index import GalleryIndex, _bits


def _pack(keys: Sequence[str]) -> Tuple[bytes, np.ndarray]:
    enc = [k.encode("utf-8") for k in keys]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        off[1:] = np.cumsum([len(e) for e in enc])
    return b"".join(enc), off


class IdMap:
    """uuid -> gallery row (`kemr_idmap_*`, host C++).  A repeated uuid keeps its last row, unknown keys map to -1,
    URIs are cut to their last '/' segment (`fusion.py:62,76-78`)."""

    def __init__(self, uuids: Sequence[str]):
        self.uuids = list(uuids)
        blob, off = _pack(self.uuids)
        h = C.c_void_p()
        self._lib = _lib.load()
        _lib.check(self._lib.kemr_idmap_create(blob, off.ctypes.data_as(C.c_void_p), len(self.uuids), C.byref(h)))
        self._h = h

    def __len__(self):
        return len(self.uuids)

    def rows(self, keys: Sequence[str], normalize_uri: bool = True) -> np.ndarray:
        blob, off = _pack(keys)
        out = np.empty(len(keys), dtype=np.int64)
        _lib.check(self._lib.kemr_idmap_lookup(self._h, blob, off.ctypes.data_as(C.c_void_p), len(keys),
                                               int(normalize_uri), out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kemr_idmap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_text2sparql_results(directory: str) -> Dict[str, List[str]]:
    """`evaluator.py:43-50`: one file per query, named `<query uuid>.<ext>`, one artefact URI per line."""
    out: Dict[str, List[str]] = {}
    for name in sorted(os.listdir(directory)):
        with open(os.path.join(directory, name), "r") as f:
            out[name.split(".")[0]] = [line.strip() for line in f.readlines()]
    return out


class HitLists:
    """Per-query KG result lists mapped to GLOBAL gallery rows, resident on the device, from which the CSR of any
    shard and any fusion strategy is built by `kemr_hits_build_csr` without going back to strings."""

    def __init__(self, idmap: IdMap, results: Dict[str, List[str]], query_uuids: Sequence[str]):
        lists = [results.get(qu, []) for qu in query_uuids]
        self.sizes = np.array([len(l) for l in lists], dtype=np.int64)            # raw list lengths (fusion.py:185)
        rowptr = np.zeros(len(lists) + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(self.sizes)
        flat = [u for l in lists for u in l]
        rows = idmap.rows(flat) if flat else np.empty(0, dtype=np.int64)
        self.Q = len(lists)
        self.nnz = int(rowptr[-1])
        self.rowptr = torch.from_numpy(rowptr).cuda()
        self.rows = torch.from_numpy(rows).cuda() if self.nnz else torch.empty(0, dtype=torch.int64, device="cuda")

    def csr(self, bonus_per_query, sum_repeats: bool, row_lo: int = 0, row_hi: Optional[int] = None) -> engine.KGHits:
        """CSR of the shard [row_lo, row_hi) with unique local columns; `bonus_per_query` is a scalar or [Q] array:
        the bonus of one listing (w, delta, delta*omega)."""
        lib = _lib.load()
        hi = (1 << 31) - 1 + row_lo if row_hi is None else row_hi
        b = torch.as_tensor(np.broadcast_to(np.asarray(bonus_per_query, dtype=np.float64), (self.Q,)).copy()).cuda()
        rowptr = torch.empty(self.Q + 1, dtype=torch.int64, device="cuda")
        col = torch.empty(max(1, self.nnz), dtype=torch.int32, device="cuda")
        bonus = torch.empty(max(1, self.nnz), dtype=torch.float64, device="cuda")
        mx = torch.zeros(1, dtype=torch.int64, device="cuda")
        ws = torch.empty(int(lib.kemr_hits_workspace_bytes(self.Q)), dtype=torch.uint8, device="cuda")
        p = engine._ptr
        _lib.check(lib.kemr_hits_build_csr(p(self.rowptr), p(self.rows), p(b), self.Q, int(row_lo), int(hi),
                                           int(bool(sum_repeats)), p(rowptr), p(col), p(bonus), p(mx), p(ws), ws.numel(),
                                           engine._stream()))
        n = int(rowptr[-1].item())
        return engine.KGHits(rowptr, col[:n], bonus[:n], int(mx.item()))

    def for_strategy(self, fusion_strategy: str = "weighted", fusion_params: Optional[Dict] = None, row_lo: int = 0,
                     row_hi: Optional[int] = None) -> Tuple[float, engine.KGHits]:
        """(alpha, CSR) reproducing `fuse_clip_and_text2sparql` (`fusion.py:209-276`) as final = alpha*clip + bonus."""
        from .fusion import _DEFAULT_OMEGA, _omega
        p = fusion_params or {}
        if fusion_strategy == "weighted":
            alpha, w = p.get("alpha", 0.7), p.get("sparql_weight", 0.3)
            if not np.isclose(alpha + w, 1.0):
                tot = alpha + w
                alpha, w = alpha / tot, w / tot
            return alpha, self.csr(w, False, row_lo, row_hi)
        if fusion_strategy == "additive":
            return 1.0, self.csr(p.get("delta", 0.5), True, row_lo, row_hi)
        if fusion_strategy == "adaptive":
            d = p.get("delta", 0.5)
            th = p.get("size_thresholds") or dict(_DEFAULT_OMEGA)
            return 1.0, self.csr(np.array([d * _omega(int(n), th) for n in self.sizes]), True, row_lo, row_hi)
        raise ValueError(f"Unknown fusion strategy: {fusion_strategy}")


class EmbeddingStore:
    """A gallery persisted as bf16 shardable files plus its uuid list."""

    def __init__(self, directory: str):
        self.dir = directory
        lib = _lib.load()
        rows, dim = C.c_int64(), C.c_int()
        _lib.check(lib.kemr_store_info(self._path("image").encode(), C.byref(rows), C.byref(dim)))
        self.M, self.D = rows.value, dim.value
        self.has_target = os.path.exists(self._path("target"))
        with open(os.path.join(directory, "uuids.txt")) as f:
            self.uuids = [line.rstrip("\n") for line in f]
        if len(self.uuids) != self.M:
            raise KemrError(f"{directory}: {len(self.uuids)} uuids for {self.M} rows")
        self._idmap: Optional[IdMap] = None

    def _path(self, name: str) -> str:
        return os.path.join(self.dir, f"{name}.kemr")

    @staticmethod
    def save(directory: str, image_embeddings, target_embeddings, uuids: Sequence[str]) -> "EmbeddingStore":
        """fp32 (bf16-representable or to be rounded on upload elsewhere) or uint16 bf16 bit patterns -> files."""
        os.makedirs(directory, exist_ok=True)
        lib = _lib.load()
        for name, emb in (("image", image_embeddings), ("target", target_embeddings)):
            if emb is None:
                continue
            bits = _bits(emb)
            if bits.shape[0] != len(uuids):
                raise KemrError("embeddings rows != uuids")
            _lib.check(lib.kemr_store_write(os.path.join(directory, f"{name}.kemr").encode(),
                                            bits.ctypes.data_as(C.c_void_p), bits.shape[0], bits.shape[1]))
        with open(os.path.join(directory, "uuids.txt"), "w") as f:
            f.write("".join(u + "\n" for u in uuids))
        return EmbeddingStore(directory)

    @property
    def idmap(self) -> IdMap:
        if self._idmap is None:
            self._idmap = IdMap(self.uuids)
        return self._idmap

    def load_rows(self, name: str, lo: int, hi: int) -> torch.Tensor:
        out = torch.empty((hi - lo, self.D), dtype=torch.bfloat16, device="cuda")
        if hi > lo:
            _lib.check(_lib.load().kemr_store_load(self._path(name).encode(), lo, hi, engine._ptr(out), engine._stream()))
        return out

    def load(self, rank: int = 0, world: int = 1) -> GalleryIndex:
        """This rank's contiguous row shard as a device-resident GalleryIndex emitting GLOBAL row ids."""
        lo, hi = shard_bounds(self.M, world, rank)
        img = self.load_rows("image", lo, hi)
        tgt = self.load_rows("target", lo, hi) if self.has_target else None
        return GalleryIndex(img, tgt, self.uuids[lo:hi], idx_base=lo)
