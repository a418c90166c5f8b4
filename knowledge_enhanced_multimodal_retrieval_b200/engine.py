"""Device-level Python API over libkemr.so.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed only; every
computation on the hot path is a kernel of libkemr.so reached through the C ABI of
include/kemr.h.  All functions take/return CUDA tensors and enqueue on the current stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from ._lib import KemrError, PATH_AUTO, PATH_MMA, PATH_WARP, FLAG_OVERFLOW, FLAG_UNCERTIFIED  # noqa: F401  (PATH_* re-exported)

# Bound on |fp32 scan score - canonical binary64 score| for ||q||, ||g|| <= 1 and |w_a| + |w_b| <= 1:
# every accumulator update rounds (or truncates) once, relative to a running magnitude of at most sum_d |q_d g_d|
# <= ||q|| ||g||, so |error| <= 2^-23 * (D_acc / 16 + 8) * S with D_acc the accumulated length (2 D when two galleries
# share an accumulator) and S = max||q|| * max||g|| * (|w_a| + |w_b|); 1.25 * 2^-23 * (2048 / 16 + 8) = 2.03e-5 at the
# largest supported shape (D = 1024, two galleries).  Un-normalised inputs scale it: see `eps_for`.
DEFAULT_EPS = 2e-5
MAX_K_SEL = 128

ArrayLike = Union[np.ndarray, torch.Tensor]


def _require_cuda():
    if not torch.cuda.is_available():
        raise KemrError("no CUDA device: this engine has no CPU path")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def device_info() -> Dict[str, int]:
    _require_cuda()
    lib = _lib.load()
    v = [C.c_int() for _ in range(4)]
    _lib.check(lib.kemr_device_info(*[C.byref(x) for x in v]))
    return {"sm_count": v[0].value, "cc_major": v[1].value, "cc_minor": v[2].value,
            "has_tcgen05": v[3].value}


def scan_plan(Q: int, M: int, D: int, galleries: int = 1, k_sel: int = 16, equal_weights: bool = False) -> Dict[str, int]:
    """Kernel KEMR_PATH_AUTO would run for this shape (PATH_WARP / PATH_MMA) and its part count."""
    path, parts = C.c_int(), C.c_int()
    _lib.check(_lib.load().kemr_scan_plan(int(Q), int(M), int(D), int(galleries), int(k_sel), int(bool(equal_weights)),
                                          C.byref(path), C.byref(parts)))
    return {"path": path.value, "parts": parts.value}


# --------------------------------------------------------------------------- workspace
class _Workspace:
    """One scratch buffer per device, grown on demand and reused by every call of this module.  Calls are enqueued
    on the caller's current stream and share this buffer, so two scans must not be in flight on DIFFERENT streams at
    the same time (the reference is single-threaded); callers that want that pass their own workspace to
    `scan_topk_raw`."""

    def __init__(self):
        self.buf: Dict[int, torch.Tensor] = {}

    def get(self, nbytes: int) -> torch.Tensor:
        dev = torch.cuda.current_device()
        b = self.buf.get(dev)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes), dtype=torch.uint8, device=f"cuda:{dev}")
            self.buf[dev] = b
        return b


_WS = _Workspace()


def workspace_for(Q: int, M: int, D: int, k_sel: int, max_hits: int = 0, scale: int = 1) -> torch.Tensor:
    n = _lib.load().kemr_workspace_bytes(int(Q), int(M), int(D), int(k_sel), int(max_hits))
    return _WS.get(int(n) * scale)


# --------------------------------------------------------------------------- embeddings
def quantize(x: ArrayLike, normalize: bool = False) -> torch.Tensor:
    """fp32 rows (numpy / torch, host or device) -> bf16 CUDA tensor via kemr_quantize_rows.
    bf16 CUDA tensors pass through untouched."""
    _require_cuda()
    if isinstance(x, torch.Tensor) and x.dtype == torch.bfloat16:
        if normalize:
            raise KemrError("bf16 input cannot be re-normalised")
        return x.cuda().contiguous()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    x = x.to(dtype=torch.float32)
    if x.dim() != 2:
        raise KemrError(f"embeddings must be 2-D, got shape {tuple(x.shape)}")
    xd = x.cuda(non_blocking=True).contiguous()
    out = torch.empty(xd.shape, dtype=torch.bfloat16, device=xd.device)
    if xd.shape[0]:
        _lib.check(_lib.load().kemr_quantize_rows(_ptr(xd), _ptr(out), xd.shape[0], xd.shape[1],
                                                  int(normalize), _stream()))
    return out


def row_norm_max(x: torch.Tensor) -> float:
    """Largest Euclidean row norm of a bf16 CUDA matrix (kemr_row_norm_max), cached on the tensor object: galleries
    are measured once, a query batch once per call (one 4-byte read-back)."""
    cached = getattr(x, "_kemr_norm_max", None)
    if cached is not None:
        return cached
    out = torch.zeros(1, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().kemr_row_norm_max(_ptr(x), x.shape[0], x.shape[1], _ptr(out), _stream()))
    v = float(out.item())
    try:
        x._kemr_norm_max = v
    except Exception:            # noqa: BLE001  (tensor subclasses without a __dict__)
        pass
    return v


def eps_for(q: torch.Tensor, gal_a: torch.Tensor, gal_b: Optional[torch.Tensor] = None, w_a=1.0, w_b=0.0,
            base: float = DEFAULT_EPS) -> float:
    """Selection margin for THESE embeddings and weights: DEFAULT_EPS * max(1, S), S = max||q|| * (|w_a| max||g_a|| +
    |w_b| max||g_b||) (per-query weights: their largest magnitudes).  L2-normalised embeddings with convex weights give
    S <= 1 up to bf16 rounding, i.e. the default; projected galleries (BilinearFusionHead), raw features or weights
    above one widen the margin instead of silently voiding the certificate."""
    def wmax(w):
        if isinstance(w, torch.Tensor):
            return float(w.abs().max().item()) if w.numel() else 0.0
        if isinstance(w, np.ndarray):
            return float(np.abs(w).max()) if w.size else 0.0
        return abs(float(w))
    s = row_norm_max(gal_a) * wmax(w_a)
    if gal_b is not None:
        s += row_norm_max(gal_b) * wmax(w_b)
    s *= row_norm_max(q)
    if not np.isfinite(s):
        raise KemrError("embeddings contain non-finite values: the scan paths need finite inputs")
    return base * max(1.0, s * (1.0 + 2.0 ** -7))


def synth_rows(rows: int, D: int, seed: int, row_base: int = 0, out: Optional[torch.Tensor] = None):
    """Deterministic synthetic gallery shard generated on the device (kemr_synth_rows)."""
    _require_cuda()
    if out is None:
        out = torch.empty((rows, D), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.load().kemr_synth_rows(_ptr(out), rows, D, C.c_uint64(seed), row_base, _stream()))
    return out


# --------------------------------------------------------------------------- KG hits
@dataclass
class KGHits:
    """CSR of knowledge-graph hits per query: unique local gallery columns + aggregated bonus."""
    rowptr: torch.Tensor      # int64 [Q+1]
    col: torch.Tensor         # int32 [nnz]
    bonus: torch.Tensor       # float64 [nnz]
    max_per_query: int

    @staticmethod
    def from_lists(cols_per_query: Sequence[Sequence[int]], bonus_per_query: Sequence[Sequence[float]],
                   device="cuda") -> "KGHits":
        rowptr = np.zeros(len(cols_per_query) + 1, dtype=np.int64)
        for i, c in enumerate(cols_per_query):
            rowptr[i + 1] = rowptr[i] + len(c)
        col = np.fromiter((c for cs in cols_per_query for c in cs), dtype=np.int32, count=int(rowptr[-1]))
        bon = np.fromiter((b for bs in bonus_per_query for b in bs), dtype=np.float64, count=int(rowptr[-1]))
        mx = int(np.diff(rowptr).max()) if len(cols_per_query) else 0
        return KGHits(torch.from_numpy(rowptr).to(device), torch.from_numpy(col).to(device),
                      torch.from_numpy(bon).to(device), mx)

    def _filter(self, sel: Optional[torch.Tensor], n_out: int, lo: int, hi: int) -> "KGHits":
        """kemr_hits_filter_csr: query subset and / or column range, on the device."""
        lib = _lib.load()
        dev = self.rowptr.device
        nnz = self.col.numel()
        rowptr = torch.empty(n_out + 1, dtype=torch.int64, device=dev)
        col = torch.empty(max(1, nnz), dtype=torch.int32, device=dev)
        bonus = torch.empty(max(1, nnz), dtype=torch.float64, device=dev)
        mx = torch.zeros(1, dtype=torch.int64, device=dev)
        ws = torch.empty(int(lib.kemr_hits_workspace_bytes(n_out)), dtype=torch.uint8, device=dev)
        _lib.check(lib.kemr_hits_filter_csr(_ptr(self.rowptr), _ptr(self.col), _ptr(self.bonus), _ptr(sel), n_out, int(lo), int(hi),
                                            _ptr(rowptr), _ptr(col), _ptr(bonus), _ptr(mx), _ptr(ws), ws.numel(), _stream()))
        tail = torch.stack([rowptr[-1], mx[0]]).cpu()          # one read-back: nnz and the longest list
        n = int(tail[0])
        return KGHits(rowptr, col[:n], bonus[:n], int(tail[1]))

    def subset(self, rows: torch.Tensor) -> "KGHits":
        """CSR restricted to the given query rows (in that order)."""
        sel = rows.to(device=self.rowptr.device, dtype=torch.int64).contiguous()
        if sel.numel() == 0:
            return KGHits(torch.zeros(1, dtype=torch.int64, device=self.rowptr.device), self.col[:0], self.bonus[:0], 0)
        return self._filter(sel, sel.numel(), 0, 1 << 31)

    def shard(self, lo: int, hi: int) -> "KGHits":
        """Hits falling in gallery rows [lo, hi), re-based to local indices."""
        return self._filter(None, self.rowptr.numel() - 1, lo, hi)


def uri_tail(uri: str) -> str:
    """Last '/' segment of an artefact URI (reference: fusion.py:76)."""
    return uri.rsplit("/", 1)[-1] if "/" in uri else uri


def kg_pairs(results: Dict[str, List[str]], query_uuids: Sequence[str], artefact_uuids: Sequence[str]):
    """Per query: gallery columns of the known hits in list order, and the raw list length.
    Unknown queries/artefacts contribute nothing (fusion.py:70,78)."""
    col = {u: j for j, u in enumerate(artefact_uuids)}
    cols, sizes = [], []
    for qu in query_uuids:
        lst = results.get(qu, [])
        sizes.append(len(lst))
        cs = []
        for uri in lst:
            j = col.get(uri_tail(uri))
            if j is not None:
                cs.append(j)
        cols.append(cs)
    return cols, sizes


def build_hits(results, query_uuids, artefact_uuids, bonus_of_query, dedupe: bool) -> KGHits:
    """CSR with unique columns per query.  `bonus_of_query(i, list_len)` is the binary64 bonus
    of ONE listing; with dedupe=False repeated listings add up (fusion.py:130)."""
    cols, sizes = kg_pairs(results, query_uuids, artefact_uuids)
    out_c, out_b = [], []
    for i, cs in enumerate(cols):
        b = float(bonus_of_query(i, sizes[i])) if cs else 0.0
        agg: Dict[int, float] = {}
        for c in cs:
            agg[c] = b if (dedupe or c not in agg) else agg[c] + b
        out_c.append(list(agg.keys()))
        out_b.append(list(agg.values()))
    return KGHits.from_lists(out_c, out_b)


# --------------------------------------------------------------------------- scan + top-k
def default_k_sel(k: int) -> int:
    """k plus a margin of >= 6, rounded up to the register-list sizes of the tcgen05 epilogue."""
    return min(MAX_K_SEL, (k + 6 + 7) // 8 * 8)


def _per_query(w) -> bool:
    return isinstance(w, (torch.Tensor, np.ndarray))


def query_weights(w_a, w_b, Q: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-query fusion weights (gated heads) as contiguous binary64 device arrays [Q]."""
    out = []
    for w in (w_a, w_b):
        t = torch.as_tensor(w).to(device=device, dtype=torch.float64).reshape(-1).contiguous()
        if t.numel() != Q:
            raise KemrError(f"per-query weights need {Q} entries, got {t.numel()}")
        out.append(t)
    return out[0], out[1]


def gate_linear(q: torch.Tensor, weight: ArrayLike, bias: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """(gate, 1 - gate) of the linear gated heads, gate = sigmoid(q . weight + bias) in fp32
    (reference `fusion_model.py:18-19,190-191`), as binary64 per-query weights for scan_topk / rank_targets."""
    Q, D = q.shape
    w = torch.as_tensor(weight).to(device=q.device, dtype=torch.float32).reshape(-1).contiguous()
    if w.numel() != D:
        raise KemrError(f"gate weight needs {D} entries, got {w.numel()}")
    wa = torch.empty((Q,), dtype=torch.float64, device=q.device)
    wb = torch.empty((Q,), dtype=torch.float64, device=q.device)
    _lib.check(_lib.load().kemr_gate_linear(_ptr(q), Q, D, _ptr(w), float(bias), _ptr(wa), _ptr(wb), _stream()))
    return wa, wb


def scan_topk_raw(q, gal_a, gal_b, w_a, w_b, alpha, hits: Optional[KGHits], k, k_sel, eps, idx_base,
                  out_score, out_idx, out_flags, ws, path=PATH_AUTO, out_score32=None):
    """One kemr_scan_topk call (kemr_scan_topk_gated when w_a / w_b are per-query arrays); no host synchronisation."""
    Q, D = q.shape
    M = gal_a.shape[0]
    if _per_query(w_a) or _per_query(w_b):
        wa, wb = query_weights(w_a, w_b, Q, q.device)
        _lib.check(_lib.load().kemr_scan_topk_gated(
            _ptr(q), Q, _ptr(gal_a), _ptr(gal_b), M, D, _ptr(wa), _ptr(wb), float(alpha),
            _ptr(hits.rowptr) if hits else None, _ptr(hits.col) if hits else None,
            _ptr(hits.bonus) if hits else None, hits.max_per_query if hits else 0,
            k, k_sel, float(eps), int(idx_base), _ptr(out_score), _ptr(out_score32), _ptr(out_idx),
            _ptr(out_flags), _ptr(ws), ws.numel(), path, _stream()))
        return
    _lib.check(_lib.load().kemr_scan_topk(
        _ptr(q), Q, _ptr(gal_a), _ptr(gal_b), M, D, float(w_a), float(w_b), float(alpha),
        _ptr(hits.rowptr) if hits else None, _ptr(hits.col) if hits else None,
        _ptr(hits.bonus) if hits else None, hits.max_per_query if hits else 0,
        k, k_sel, float(eps), int(idx_base), _ptr(out_score), _ptr(out_score32), _ptr(out_idx),
        _ptr(out_flags), _ptr(ws), ws.numel(), path, _stream()))


def scan_topk(q: torch.Tensor, gal_a: torch.Tensor, gal_b: Optional[torch.Tensor] = None,
              w_a: float = 1.0, w_b: float = 0.0, alpha: float = 1.0, hits: Optional[KGHits] = None,
              k: int = 10, k_sel: Optional[int] = None, eps: Optional[float] = None, idx_base: int = 0,
              path: int = PATH_AUTO, certify: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fused similarity scan + weighted fusion + KG boost + top-k.

    Returns (idx int64 [Q,k], score float64 [Q,k]) ordered by (canonical score desc, index asc).
    With certify=True the per-query certificate is read back and uncertified queries are re-run
    with a wider selection margin (raises if the margin cannot be certified at k_sel=128).
    eps=None scales the selection margin with the embeddings' norms and the weights (`eps_for`; one small
    read-back per new query batch); pass a float to skip that (DEFAULT_EPS holds for normalised embeddings and
    convex weights).
    """
    _require_cuda()
    _check_pair(q, gal_a, gal_b)
    Q, D = q.shape
    M = gal_a.shape[0]
    k_sel = default_k_sel(k) if k_sel is None else k_sel
    mh = hits.max_per_query if hits else 0
    if eps is None:
        eps = eps_for(q, gal_a, gal_b, w_a, w_b)
    gated = _per_query(w_a) or _per_query(w_b)               # per-query weights (gated fusion heads)
    if gated:
        w_a, w_b = query_weights(w_a, w_b, Q, q.device)
    ws = workspace_for(Q, M, D, k_sel, mh)
    score = torch.empty((Q, k), dtype=torch.float64, device=q.device)
    idx = torch.empty((Q, k), dtype=torch.int64, device=q.device)
    flags = torch.empty((Q,), dtype=torch.int32, device=q.device)
    scan_topk_raw(q, gal_a, gal_b, w_a, w_b, alpha, hits, k, k_sel, eps, idx_base, score, idx, flags, ws, path)
    if certify:
        bad = torch.nonzero(flags & FLAG_UNCERTIFIED).flatten()
        ks = k_sel
        while bad.numel():
            if ks >= MAX_K_SEL:
                # more rows than the widest selection holds sit within 2*eps of the k-th score (duplicated or
                # degenerate embeddings): decide those queries exactly instead of giving up (the reference just
                # returns a ranking)
                _exact_topk(q, gal_a, gal_b, w_a, w_b, alpha, hits, k, eps, idx_base, bad, gated, score, idx, path)
                flags[bad] = 0
                break
            ks = min(MAX_K_SEL, ks * 2)
            sub_hits = hits.subset(bad) if hits is not None else None
            qs = q[bad].contiguous()
            n = qs.shape[0]
            wa_sub, wb_sub = (w_a[bad].contiguous(), w_b[bad].contiguous()) if gated else (w_a, w_b)
            s2 = torch.empty((n, k), dtype=torch.float64, device=q.device)
            i2 = torch.empty((n, k), dtype=torch.int64, device=q.device)
            f2 = torch.empty((n,), dtype=torch.int32, device=q.device)
            ws2 = workspace_for(n, M, D, ks, sub_hits.max_per_query if sub_hits else 0)
            scan_topk_raw(qs, gal_a, gal_b, wa_sub, wb_sub, alpha, sub_hits, k, ks, eps, idx_base, s2, i2, f2, ws2, path)
            score[bad] = s2
            idx[bad] = i2
            flags[bad] = f2
            bad = bad[torch.nonzero(f2 & FLAG_UNCERTIFIED).flatten()]
    _LAST_FLAGS[0] = flags
    return idx, score


def _exact_topk(q, gal_a, gal_b, w_a, w_b, alpha, hits, k, eps, idx_base, bad, gated, score, idx, path):
    """Exact top-k of the queries `bad` whatever the data: every row whose fp32 scan score lies within 2*eps of the
    k-th best fp32 score (plus every KG hit) is re-scored canonically and ordered by (score desc, index asc).  Slow
    path for degenerate inputs only (hundreds of rows tied around the k-th score)."""
    M = gal_a.shape[0]
    for qi in bad.tolist():
        qs = q[qi:qi + 1].contiguous()
        wa1, wb1 = (w_a[qi:qi + 1].contiguous(), w_b[qi:qi + 1].contiguous()) if gated else (w_a, w_b)
        wa32, wb32 = (float(wa1[0]), float(wb1[0])) if gated else (float(w_a), float(w_b))
        dense = score_matrix(qs, gal_a, gal_b, wa32, wb32, path=path)[0]
        kk = min(k, M)
        thr = torch.topk(dense, kk).values[-1] - 2.0 * eps * (1.0 + 1.0 / 64.0) - 1e-30
        rows = torch.nonzero(dense >= thr).flatten()
        bonus = None
        if hits is not None:
            h0, h1 = int(hits.rowptr[qi]), int(hits.rowptr[qi + 1])
            hc = hits.col[h0:h1].to(torch.int64)
            rows = torch.unique(torch.cat([rows, hc[(hc >= 0) & (hc < M)]]))
            b = torch.zeros(M, dtype=torch.float64, device=q.device)
            b[hc[(hc >= 0) & (hc < M)]] = hits.bonus[h0:h1][(hc >= 0) & (hc < M)]
            bonus = b[rows]
        sc = score_pairs(qs, gal_a, gal_b, torch.zeros_like(rows, dtype=torch.int32), rows, wa1 if gated else w_a,
                         wb1 if gated else w_b, alpha, bonus)
        order = torch.argsort(rows)                                    # rows ascending, then a STABLE sort by score desc
        rows, sc = rows[order], sc[order]
        order = torch.sort(-sc, stable=True).indices[:kk]
        score[qi, :kk] = sc[order]
        idx[qi, :kk] = rows[order] + idx_base
        if kk < k:
            score[qi, kk:] = float("-inf")
            idx[qi, kk:] = -1


_LAST_FLAGS: List[Optional[torch.Tensor]] = [None]


def last_flags() -> Optional[torch.Tensor]:
    """Certificate flags of the most recent scan_topk call (int32 [Q])."""
    return _LAST_FLAGS[0]


def _check_pair(q, gal_a, gal_b):
    for t in (q, gal_a) + ((gal_b,) if gal_b is not None else ()):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2
                and t.is_contiguous()):
            raise KemrError("expected contiguous 2-D bf16 CUDA tensors (use engine.quantize)")
    if q.shape[1] != gal_a.shape[1] or (gal_b is not None and gal_b.shape != gal_a.shape):
        raise KemrError(f"shape mismatch: q {tuple(q.shape)}, gal_a {tuple(gal_a.shape)}, "
                        f"gal_b {None if gal_b is None else tuple(gal_b.shape)}")


def score_pairs(q, gal_a, gal_b, pair_q: torch.Tensor, pair_row: torch.Tensor, w_a=1.0, w_b=0.0, alpha=1.0,
                pair_bonus: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Canonical binary64 final score of (query, row) pairs."""
    _check_pair(q, gal_a, gal_b)
    n = pair_q.numel()
    out = torch.empty((n,), dtype=torch.float64, device=q.device)
    pq = pair_q.to(device=q.device, dtype=torch.int32).contiguous()
    pr = pair_row.to(device=q.device, dtype=torch.int64).contiguous()
    pb = None if pair_bonus is None else pair_bonus.to(device=q.device, dtype=torch.float64).contiguous()
    if _per_query(w_a) or _per_query(w_b):
        wa, wb = query_weights(w_a, w_b, q.shape[0], q.device)
        _lib.check(_lib.load().kemr_score_pairs_gated(_ptr(q), _ptr(gal_a), _ptr(gal_b), gal_a.shape[0], q.shape[1], _ptr(wa), _ptr(wb),
                                                      float(alpha), _ptr(pq), _ptr(pr), _ptr(pb), n, _ptr(out), _stream()))
        return out
    _lib.check(_lib.load().kemr_score_pairs(_ptr(q), _ptr(gal_a), _ptr(gal_b), gal_a.shape[0], q.shape[1], float(w_a),
                                            float(w_b), float(alpha), _ptr(pq), _ptr(pr), _ptr(pb), n,
                                            _ptr(out), _stream()))
    return out


def target_bonus(hits: Optional[KGHits], target_idx: torch.Tensor) -> Optional[torch.Tensor]:
    """Bonus of each query's own target row (0 if it is not a KG hit of that query), on the device."""
    if hits is None:
        return None
    t = target_idx.to(device=hits.rowptr.device, dtype=torch.int64).contiguous()
    out = torch.empty(t.numel(), dtype=torch.float64, device=t.device)
    if t.numel():
        _lib.check(_lib.load().kemr_hits_target_bonus(_ptr(hits.rowptr), _ptr(hits.col), _ptr(hits.bonus), t.numel(),
                                                      _ptr(t), _ptr(out), _stream()))
    return out


def rank_count(q, gal_a, gal_b, t_score: torch.Tensor, t_gidx: torch.Tensor, w_a=1.0, w_b=0.0, alpha=1.0,
               hits: Optional[KGHits] = None, eps: Optional[float] = None, idx_base: int = 0,
               path: int = PATH_AUTO) -> torch.Tensor:
    """#rows of this shard ranked strictly ahead of each query's target (int64 [Q]).  eps=None: `eps_for`."""
    _check_pair(q, gal_a, gal_b)
    if eps is None:
        eps = eps_for(q, gal_a, gal_b, w_a, w_b)
    Q, D = q.shape
    M = gal_a.shape[0]
    scale = 1
    while True:
        ws = workspace_for(Q, M, D, 16, 0, scale)
        count = torch.empty((Q,), dtype=torch.int64, device=q.device)
        flags = torch.empty((Q,), dtype=torch.int32, device=q.device)
        tail = (_ptr(hits.rowptr) if hits else None, _ptr(hits.col) if hits else None,
                _ptr(hits.bonus) if hits else None, _ptr(t_score), _ptr(t_gidx), float(eps), int(idx_base),
                _ptr(count), _ptr(flags), _ptr(ws), ws.numel(), path, _stream())
        if _per_query(w_a) or _per_query(w_b):
            wa, wb = query_weights(w_a, w_b, Q, q.device)
            _lib.check(_lib.load().kemr_rank_count_gated(_ptr(q), Q, _ptr(gal_a), _ptr(gal_b), M, D, _ptr(wa), _ptr(wb),
                                                         float(alpha), *tail))
        else:
            _lib.check(_lib.load().kemr_rank_count(_ptr(q), Q, _ptr(gal_a), _ptr(gal_b), M, D, float(w_a), float(w_b),
                                                   float(alpha), *tail))
        if not bool((flags & FLAG_OVERFLOW).any()):
            return count
        if scale >= 64:
            raise KemrError("ambiguous-candidate list keeps overflowing; eps is too large for this gallery")
        scale *= 4


def rank_targets(q, gal_a, gal_b, target_idx: torch.Tensor, w_a=1.0, w_b=0.0, alpha=1.0,
                 hits: Optional[KGHits] = None, eps: Optional[float] = None, path: int = PATH_AUTO) -> torch.Tensor:
    """1-based rank of gallery row target_idx[i] for query i under the canonical ordering."""
    Q = q.shape[0]
    tidx = target_idx.to(device=q.device, dtype=torch.int64).contiguous()
    t = score_pairs(q, gal_a, gal_b, torch.arange(Q, device=q.device), tidx, w_a, w_b, alpha,
                    target_bonus(hits, tidx))
    return rank_count(q, gal_a, gal_b, t, tidx, w_a, w_b, alpha, hits, eps, 0, path) + 1


def score_matrix(q, gal_a, gal_b=None, w_a=1.0, w_b=0.0, path: int = PATH_AUTO) -> torch.Tensor:
    """Dense fp32 fused similarity matrix [Q, M] as the scan kernels compute it."""
    _check_pair(q, gal_a, gal_b)
    Q, D = q.shape
    M = gal_a.shape[0]
    out = torch.empty((Q, M), dtype=torch.float32, device=q.device)
    ws = workspace_for(Q, M, D, 16)
    _lib.check(_lib.load().kemr_score_matrix(_ptr(q), Q, _ptr(gal_a), _ptr(gal_b), M, D, float(w_a), float(w_b),
                                             _ptr(out), M, _ptr(ws), ws.numel(), path, _stream()))
    return out


# --------------------------------------------------------------------------- matrix-taking ops
def _as_f32_matrix(S: ArrayLike) -> torch.Tensor:
    _require_cuda()
    if isinstance(S, np.ndarray):
        S = torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32))
    S = S.to(dtype=torch.float32)
    if S.dim() != 2:
        raise KemrError("similarity matrix must be 2-D")
    return S.cuda(non_blocking=True).contiguous()


def matrix_rank(S: ArrayLike, target_col: Optional[torch.Tensor] = None) -> torch.Tensor:
    Sd = _as_f32_matrix(S)
    Q, M = Sd.shape
    if target_col is None:
        target_col = torch.arange(Q, device=Sd.device, dtype=torch.int64)     # metrics.py:37
    target_col = target_col.to(device=Sd.device, dtype=torch.int64).contiguous()
    out = torch.empty((Q,), dtype=torch.int64, device=Sd.device)
    _lib.check(_lib.load().kemr_matrix_rank(_ptr(Sd), Q, M, M, _ptr(target_col), _ptr(out), _stream()))
    return out


def matrix_topk(S: ArrayLike, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    Sd = _as_f32_matrix(S)
    Q, M = Sd.shape
    idx = torch.empty((Q, k), dtype=torch.int64, device=Sd.device)
    val = torch.empty((Q, k), dtype=torch.float32, device=Sd.device)
    _lib.check(_lib.load().kemr_matrix_topk(_ptr(Sd), Q, M, M, k, _ptr(idx), _ptr(val), _stream()))
    return idx, val


def matrix_fuse(S: ArrayLike, scale_first: bool, alpha32: float, rowptr: np.ndarray, col: np.ndarray,
                add: np.ndarray) -> torch.Tensor:
    Sd = _as_f32_matrix(S)
    Q, M = Sd.shape
    out = torch.empty_like(Sd)
    rp = torch.from_numpy(np.ascontiguousarray(rowptr, dtype=np.int64)).cuda()
    cc = torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)).cuda()
    aa = torch.from_numpy(np.ascontiguousarray(add, dtype=np.float32)).cuda()
    _lib.check(_lib.load().kemr_matrix_fuse(_ptr(Sd), _ptr(out), Q, M, M, int(scale_first), float(alpha32),
                                            _ptr(rp), _ptr(cc), _ptr(aa), _stream()))
    return out


def metrics_reduce(ranks: torch.Tensor, k_values: Sequence[int]):
    """Fused Recall@K/MRR/Mean-Rank reduction on the device.
    Returns (hits int64 [n_k], sum_rank float, sum_reciprocal_rank float) on the host."""
    ranks = ranks.to(dtype=torch.int64).contiguous()
    kv = torch.tensor(list(k_values), dtype=torch.int32, device=ranks.device)
    hits = torch.zeros((max(1, len(k_values)),), dtype=torch.int64, device=ranks.device)
    stats = torch.zeros((2,), dtype=torch.float64, device=ranks.device)
    _lib.check(_lib.load().kemr_metrics_reduce(_ptr(ranks), ranks.numel(), _ptr(kv), len(k_values), _ptr(hits),
                                               _ptr(stats), _stream()))
    h = hits.cpu().numpy()[:len(k_values)]
    s = stats.cpu().numpy()
    return h, float(s[0]), float(s[1])


def metrics_reduce_host(ranks: np.ndarray, k_values: Sequence[int]):
    """Host twin (no GPU): same outputs as metrics_reduce."""
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    kv = np.ascontiguousarray(list(k_values), dtype=np.int32)
    hits = np.zeros(max(1, len(kv)), dtype=np.int64)
    stats = np.zeros(2, dtype=np.float64)
    _lib.check(_lib.load().kemr_metrics_reduce_host(
        r.ctypes.data_as(C.c_void_p), len(r), kv.ctypes.data_as(C.c_void_p), len(kv),
        hits.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p)))
    return hits[:len(kv)], float(stats[0]), float(stats[1])


def merge_topk(scores: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merge R per-shard lists [R, Q, k] into the global top-k (score desc, index asc)."""
    R, Q, kk = scores.shape
    assert kk == k and idx.shape == scores.shape
    scores = scores.to(dtype=torch.float64).contiguous()
    idx = idx.to(dtype=torch.int64).contiguous()
    os_ = torch.empty((Q, k), dtype=torch.float64, device=scores.device)
    oi = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    _lib.check(_lib.load().kemr_merge_topk(_ptr(scores), _ptr(idx), R, Q, k, _ptr(os_), _ptr(oi), _stream()))
    return oi, os_
