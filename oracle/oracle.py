"""CPU ORACLE for the retrieval-scoring hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product package never does: it fails loudly when the
CUDA library is missing instead of falling back to anything in here.

Two layers, both numpy:

1. ``ref_*``  -- a restatement of the reference's own algorithm (fp32 BLAS similarity,
   full-row argsort, dense KG-indicator fusion, python-float list fusion).  Each function
   cites the reference file:line it follows (paths relative to /root/reference).  These
   are pinned against the UNMODIFIED reference functions, imported in the build container
   by `tests/golden/make_golden.py`; the resulting vectors are committed under
   `tests/golden/` and re-checked on every CPU test run.

2. ``canon_*`` -- the refined, fully deterministic contract the GPU engine is held to
   bit-for-bit.  The reference leaves two things unspecified: the summation order of the
   BLAS sgemm (`metrics.py:102`) and the tie order of numpy's non-stable argsort
   (`metrics.py:34,62`).  The canonical oracle fixes both:
     * score: products of two bf16 values are exact in binary64; they are accumulated in
       binary64 in a fixed order -- 256 running sums (one per position inside a 256-element
       block, blocks in increasing order), an 8-to-1 tree inside every 16-byte piece, then a
       fixed halving tree (16,8,4,2,1) over the 32 pieces -- which is the order a 32-lane
       warp doing coalesced 16-byte loads uses;
     * fusion: ``fl(fl(w_a*S_a) + fl(w_b*S_b))`` then ``fl(fl(alpha*clip) + bonus)`` in
       binary64 with the python-float weights taken as doubles;
     * ranking: descending score, ties broken by the lowest gallery index (what a stable
       sort of the negated row yields).
   On inputs without fp32-level near-ties the two layers agree on every index and metric
   (checked in the tests); where they can differ the reference itself is not reproducible
   across BLAS builds.

Parity status: the reference ships no tests, fixtures or golden vectors, and
`CLIPRetriever.search` is remote code absent from the tree (`src/clip/clip_retrieval.py:15-23`)
-- for that single entry point parity is UNPINNED; everything else is pinned by running the
reference's own functions.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

DEFAULT_K = (1, 5, 10, 20)          # metrics.py:15
DEFAULT_OMEGA = ((1, 1.0), (5, 0.8), (20, 0.5), (50, 0.3), (float("inf"), 0.1))  # fusion.py:164-170


# =========================================================================== ref layer
def ref_similarity(query: np.ndarray, cand: np.ndarray) -> np.ndarray:
    """metrics.py:102 -- fp32 `Q @ C.T` through whatever BLAS numpy links."""
    return query @ cand.T


def ref_fused_similarity(query, target, image, t2i_weight=0.5, t2t_weight=0.5) -> np.ndarray:
    """metrics.py:145-148 / evaluator.py:167-169 -- two sgemms, fp32 weighted sum."""
    return (t2i_weight * (query @ image.T)) + (t2t_weight * (query @ target.T))


def _ref_order(sim: np.ndarray, sort_kind="stable") -> np.ndarray:
    # metrics.py:34,62 use the default (non-stable) kind; the stable kind is one of the
    # orders that call may legally return and is the one the contract fixes.  TIMING legs
    # (bench.py's cpu_baseline / --impl reference) pass sort_kind=None: the reference's own
    # call, numpy's default introsort (AVX-512 vectorised on x86), which is ~4x faster than the
    # stable kind on fp32 rows -- timing the stable kind would understate the reference.
    return np.argsort(-sim, axis=1, kind=sort_kind)


def ref_recall_at_k(sim: np.ndarray, k_values: Sequence[int] = DEFAULT_K, sort_kind="stable") -> Dict[str, float]:
    """metrics.py:13-44 -- target of row i is column i; hit if it is among the first k."""
    order = _ref_order(sim, sort_kind)
    want = np.arange(sim.shape[0])[:, None]
    return {f"R@{k}": np.mean((order[:, :k] == want).any(axis=1)) * 100.0 for k in k_values}


def ref_mrr_and_mean_rank(sim: np.ndarray, sort_kind="stable") -> Dict[str, float]:
    """metrics.py:47-76 -- 1-based position of column i in row i's descending order."""
    order = _ref_order(sim, sort_kind)
    want = np.arange(sim.shape[0])[:, None]
    pos = np.argmax(order == want, axis=1) + 1
    return {"MRR": np.mean(1.0 / pos) * 100.0, "Mean_Rank": np.mean(pos)}


def _prefixed(prefix: str, d: Dict[str, float]) -> Dict[str, float]:
    return {(f"{prefix}_{k}" if prefix else k): v for k, v in d.items()}     # metrics.py:108-114


def ref_metrics_from_matrix(sim, prefix="", k_values=DEFAULT_K, compute_recall=True,
                            compute_mrr=True, sort_kind="stable") -> Dict[str, float]:
    """metrics.py:165-185 (and the tail of :104-116, :150-162)."""
    out: Dict[str, float] = {}
    if compute_recall:
        out.update(_prefixed(prefix, ref_recall_at_k(sim, k_values, sort_kind)))
    if compute_mrr:
        out.update(_prefixed(prefix, ref_mrr_and_mean_rank(sim, sort_kind)))
    return out


def ref_retrieval_metrics(query, cand, prefix="", k_values=DEFAULT_K, compute_recall=True,
                          compute_mrr=True, sort_kind="stable") -> Dict[str, float]:
    """metrics.py:79-116."""
    return ref_metrics_from_matrix(ref_similarity(query, cand), prefix, k_values,
                                   compute_recall, compute_mrr, sort_kind)


def ref_retrieval_metrics_final(query, target, image, prefix="", k_values=DEFAULT_K,
                                compute_recall=True, compute_mrr=True, t2i_weight=0.5,
                                t2t_weight=0.5, sort_kind="stable") -> Dict[str, float]:
    """metrics.py:119-162."""
    sim = ref_fused_similarity(query, target, image, t2i_weight, t2t_weight)
    return ref_metrics_from_matrix(sim, prefix, k_values, compute_recall, compute_mrr, sort_kind)


def ref_all_retrieval_metrics(query, target, image, k_values=DEFAULT_K,
                              tasks=("T2I", "I2T", "T2T"), compute_recall=True,
                              compute_mrr=True) -> Dict[str, float]:
    """metrics.py:188-252 -- T2I: q->img, I2T: img->tgt, T2T: q->tgt."""
    pairs = {"T2I": (query, image), "I2T": (image, target), "T2T": (query, target)}
    out: Dict[str, float] = {}
    for name in ("T2I", "I2T", "T2T"):
        if name in tasks:
            a, b = pairs[name]
            out.update(ref_retrieval_metrics(a, b, name, k_values, compute_recall, compute_mrr))
    return out


def uri_tail(uri: str) -> str:
    """fusion.py:76 -- last '/' segment when the entry is a URI."""
    return uri.rsplit("/", 1)[-1] if "/" in uri else uri


def kg_hits_to_pairs(results: Dict[str, List[str]], query_uuids: Sequence[str],
                     artefact_uuids: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(query row, gallery column, raw result-list length) for every known hit, in list order.

    Restates the double loop shared by fusion.py:68-80, :122-130 and :180-204: unknown
    query uuids contribute nothing, unknown artefact uuids are skipped, duplicates are
    kept (callers decide whether they matter).  Later duplicates of a uuid in
    `artefact_uuids` win, as in the dict comprehension at fusion.py:62.
    """
    col = {u: j for j, u in enumerate(artefact_uuids)}
    rows: List[int] = []
    cols: List[int] = []
    sizes = np.zeros(len(query_uuids), dtype=np.int64)
    for i, qu in enumerate(query_uuids):
        lst = results.get(qu, [])
        sizes[i] = len(lst)
        for uri in lst:
            j = col.get(uri_tail(uri))
            if j is not None:
                rows.append(i)
                cols.append(j)
    return np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), sizes


def ref_weighted_fusion(sim, results, query_uuids, artefact_uuids, alpha=0.7, sparql_weight=0.3):
    """fusion.py:22-85 -- alpha*S + w*I with a dense 0/1 indicator; renormalise if alpha+w != 1."""
    assert sim.shape[0] == len(query_uuids)
    assert sim.shape[1] == len(artefact_uuids)
    if not np.isclose(alpha + sparql_weight, 1.0):                  # fusion.py:55-59
        tot = alpha + sparql_weight
        alpha, sparql_weight = alpha / tot, sparql_weight / tot
    r, c, _ = kg_hits_to_pairs(results, query_uuids, artefact_uuids)
    ind = np.zeros_like(sim)
    ind[r, c] = 1.0
    return alpha * sim + sparql_weight * ind


def ref_additive_bonus_fusion(sim, results, query_uuids, artefact_uuids, delta=0.5):
    """fusion.py:88-132 -- copy, then `+= delta` once per listed hit (duplicates add again)."""
    assert sim.shape[0] == len(query_uuids)
    assert sim.shape[1] == len(artefact_uuids)
    r, c, _ = kg_hits_to_pairs(results, query_uuids, artefact_uuids)
    out = sim.copy()
    np.add.at(out, (r, c), np.float32(delta))      # python-float delta is cast to the array dtype
    return out


def omega_for_size(size: int, thresholds=None) -> float:
    """fusion.py:191-196 -- weight of the first (sorted) threshold >= size, else 0."""
    items = sorted(dict(thresholds).items()) if thresholds is not None else list(DEFAULT_OMEGA)
    for thr, w in items:
        if size <= thr:
            return w
    return 0.0


def ref_adaptive_additive_fusion(sim, results, query_uuids, artefact_uuids, delta=0.5,
                                 size_thresholds=None):
    """fusion.py:135-206 -- `+= delta*omega(len(result list))` per listed hit."""
    assert sim.shape[0] == len(query_uuids)
    assert sim.shape[1] == len(artefact_uuids)
    r, c, sizes = kg_hits_to_pairs(results, query_uuids, artefact_uuids)
    out = sim.copy()
    if len(r):
        bonus = np.array([delta * omega_for_size(int(sizes[i]), size_thresholds) for i in r],
                         dtype=np.float64).astype(np.float32)   # product in python floats, then fp32 add
        np.add.at(out, (r, c), bonus)
    return out


def ref_fuse_clip_and_text2sparql(sim, results, query_uuids, artefact_uuids,
                                  fusion_strategy="weighted", fusion_params=None):
    """fusion.py:209-276 -- strategy dispatch with the reference's defaults."""
    p = fusion_params or {}
    if fusion_strategy == "weighted":
        return ref_weighted_fusion(sim, results, query_uuids, artefact_uuids,
                                   p.get("alpha", 0.7), p.get("sparql_weight", 0.3))
    if fusion_strategy == "additive":
        return ref_additive_bonus_fusion(sim, results, query_uuids, artefact_uuids,
                                         p.get("delta", 0.5))
    if fusion_strategy == "adaptive":
        return ref_adaptive_additive_fusion(sim, results, query_uuids, artefact_uuids,
                                            p.get("delta", 0.5), p.get("size_thresholds"))
    raise ValueError(f"Unknown fusion strategy: {fusion_strategy}")


def ref_fuse_clip_sparql_linear(clip_results: List[dict], sparql_results: List[str],
                                alpha: float = 0.8, beta: float = 0.2) -> List[dict]:
    """src/retrieval.py:23-76 -- python-float alpha*clip + beta*hit, round(.,4), stable sort."""
    if not clip_results:
        return []
    known = set(sparql_results)
    fused = [{"uuid": it["uuid"],
              "score": round(alpha * it["score"] + beta * (1.0 if it["uuid"] in known else 0.0), 4)}
             for it in clip_results]
    fused.sort(key=lambda d: d["score"], reverse=True)
    return fused


def ref_threshold_filter(items: List[dict], threshold: float = 0) -> List[dict]:
    """src/retrieval.py:88-95,100-107."""
    return [{"uuid": it["uuid"], "score": it["score"]} for it in items
            if it.get("score", 0) >= threshold]


# =========================================================================== canonical layer
def canon_dot64(query: np.ndarray, gallery: np.ndarray) -> np.ndarray:
    """Canonical binary64 scores, shape (Q, M).  Inputs hold bf16-representable values.

    The row is zero-padded to a multiple of 256 elements.  256 running sums, one per position r
    inside a 256-element block, each add their products block after block (a product of two bf16
    values is exact in binary64, so add-after-multiply equals a fused multiply-add).  The 8 sums of
    one 16-byte piece (r = 8l .. 8l+7) are combined as ((0+1)+(2+3))+((4+5)+(6+7)), and the 32
    piece sums are folded 16, 8, 4, 2, 1 -- a warp doing coalesced 16-byte loads, lane l owning
    piece l of every block.
    """
    q = np.asarray(query, dtype=np.float64)
    g = np.asarray(gallery, dtype=np.float64)
    Q, D = q.shape
    M = g.shape[0]
    Dp = (D + 255) // 256 * 256
    if Dp != D:
        q = np.pad(q, ((0, 0), (0, Dp - D)))
        g = np.pad(g, ((0, 0), (0, Dp - D)))
    J = Dp // 256
    out = np.empty((Q, M), dtype=np.float64)
    g3 = g.reshape(M, J, 256)
    for i in range(Q):
        q3 = q[i].reshape(J, 256)
        acc = np.zeros((M, 256), dtype=np.float64)
        for j in range(J):
            acc += g3[:, j, :] * q3[j][None, :]
        a = acc.reshape(M, 32, 8)
        lane = ((a[:, :, 0] + a[:, :, 1]) + (a[:, :, 2] + a[:, :, 3])) + \
               ((a[:, :, 4] + a[:, :, 5]) + (a[:, :, 6] + a[:, :, 7]))
        for off in (16, 8, 4, 2, 1):
            lane = lane[:, :off] + lane[:, off:2 * off]
        out[i] = lane[:, 0]
    return out


def canon_fused64(s_a: np.ndarray, s_b: Optional[np.ndarray], w_a: float = 1.0, w_b: float = 0.0,
                  alpha: float = 1.0, bonus: Optional[np.ndarray] = None) -> np.ndarray:
    """fl(fl(alpha * fl(fl(w_a*S_a) + fl(w_b*S_b))) + bonus), all binary64.  w_a / w_b may be per-query arrays [Q]
    (gated fusion heads, fusion_model.py:21,178,194: gate*t2i + (1-gate)*t2t)."""
    wa = np.asarray(w_a, dtype=np.float64)
    wb = np.asarray(w_b, dtype=np.float64)
    wa = wa.reshape(-1, 1) if wa.ndim else wa
    wb = wb.reshape(-1, 1) if wb.ndim else wb
    clip = wa * s_a
    if s_b is not None:
        clip = clip + wb * s_b
    out = np.float64(alpha) * clip
    if bonus is not None:
        out = out + bonus
    return out


def canon_bonus_matrix(Q: int, M: int, rows, cols, values, dedupe: bool) -> np.ndarray:
    """Dense binary64 bonus from hit pairs (dedupe=True: indicator semantics, fusion.py:80)."""
    b = np.zeros((Q, M), dtype=np.float64)
    if dedupe:
        b[rows, cols] = values
    else:
        np.add.at(b, (rows, cols), values)
    return b


def canon_topk(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-k per row by (score desc, index asc); NaN last.  Returns (idx int64, score f64)."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return order.astype(np.int64), np.take_along_axis(scores, order, axis=1)


def canon_rank(scores: np.ndarray, target_idx: np.ndarray) -> np.ndarray:
    """1-based stable-descending rank of column target_idx[i] in row i (counting form)."""
    rows = np.arange(scores.shape[0])
    t = scores[rows, target_idx][:, None]
    cols = np.arange(scores.shape[1])[None, :]
    tnan = np.isnan(t)
    snan = np.isnan(scores)
    ahead = np.where(tnan, ~snan | (snan & (cols < target_idx[:, None])),
                     (scores > t) | ((scores == t) & (cols < target_idx[:, None])))
    return ahead.sum(axis=1).astype(np.int64) + 1


def metrics_from_ranks(ranks: np.ndarray, k_values: Iterable[int] = DEFAULT_K, prefix: str = "",
                       compute_recall=True, compute_mrr=True) -> Dict[str, float]:
    """The reductions of metrics.py:41-42,70-71 applied to 1-based ranks."""
    out: Dict[str, float] = {}
    if compute_recall:
        for k in k_values:
            out[f"R@{k}"] = np.mean(ranks <= k) * 100.0
    if compute_mrr:
        out["MRR"] = np.mean(1.0 / ranks) * 100.0
        out["Mean_Rank"] = np.mean(ranks)
    return _prefixed(prefix, out)


def near_tie_audit(scores: np.ndarray, target_idx: Optional[np.ndarray], k: int, tol: float = 4e-7):
    """How many rows could legally be ordered differently by an fp32 reference.

    Returns (rows whose top-(k+1) has an adjacent gap <= tol,
             rows with another column within tol of the target's score).
    """
    top = -np.sort(-scores, axis=1)[:, :k + 1]
    topk_risky = int(((top[:, :-1] - top[:, 1:]) <= tol).any(axis=1).sum())
    rank_risky = 0
    if target_idx is not None:
        t = scores[np.arange(scores.shape[0]), target_idx][:, None]
        close = np.abs(scores - t) <= tol
        rank_risky = int((close.sum(axis=1) > 1).sum())
    return topk_risky, rank_risky


def ref_ranks_fp32(query, image, target=None, t2i_weight=0.5, t2t_weight=0.5, chunk: int = 512,
                   tol: float = 4e-7) -> Tuple[np.ndarray, np.ndarray]:
    """The reference's own per-query positions at full size, in query chunks (the (N, M) matrix of C1 is 740 MB):
    fp32 BLAS similarity (metrics.py:102 / :145-148), full-row argsort of the negated row (metrics.py:34,62),
    `argmax(order == i) + 1` (metrics.py:68).  Target of row i is column i (metrics.py:37).
    Returns (pos int64 [N], near_tie bool [N]) -- near_tie[i] is the row-level form of `near_tie_audit`: another column
    lies within `tol` of the target's fp32 score, i.e. the reference's own position for that row depends on the
    summation order of its BLAS."""
    q = np.asarray(query, dtype=np.float32)
    n = q.shape[0]
    pos = np.empty(n, dtype=np.int64)
    near = np.zeros(n, dtype=bool)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        sim = ref_similarity(q[lo:hi], image) if target is None else \
            ref_fused_similarity(q[lo:hi], target, image, t2i_weight, t2t_weight)
        order = np.argsort(-sim, axis=1)                           # default kind, like the reference
        want = np.arange(lo, hi)[:, None]
        pos[lo:hi] = np.argmax(order == want, axis=1) + 1
        t = sim[np.arange(hi - lo), np.arange(lo, hi)][:, None]
        near[lo:hi] = (np.abs(sim - t) <= tol).sum(axis=1) > 1
    return pos, near


def ref_gate_linear(query: np.ndarray, weight: np.ndarray, bias: float) -> np.ndarray:
    """Reference `fusion_model.py:18-19` / `:190-191`: gate = sigmoid((q * w).sum(1) + b), fp32 (numpy restatement of
    the torch ops; the summation order of torch's reduction is not pinned, so compare with a tolerance)."""
    q = np.asarray(query, dtype=np.float32)
    logit = (q * np.asarray(weight, dtype=np.float32)).sum(axis=1, dtype=np.float32) + np.float32(bias)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-logit, dtype=np.float32))).astype(np.float32)


def ref_gated_scores(query, image, target, gate: np.ndarray) -> np.ndarray:
    """Reference `fusion_model.py:16-22`: gate * (q @ img.T) + (1 - gate) * (q @ tgt.T) in fp32, gate (N,) fp32."""
    q = np.asarray(query, dtype=np.float32)
    g = np.asarray(gate, dtype=np.float32).reshape(-1, 1)
    return g * (q @ np.asarray(image, dtype=np.float32).T) + (np.float32(1.0) - g) * (q @ np.asarray(target, dtype=np.float32).T)


def ref_grouped_metrics(query, cand, text_to_artifact, k_values: Sequence[int] = DEFAULT_K) -> Dict[str, float]:
    """`baselines/evaluate_text_models.py:171-224` (single mode): fp32 similarity, full argsort per row, position of
    the first candidate whose artefact is the query's own index."""
    sim = np.asarray(query, dtype=np.float32) @ np.asarray(cand, dtype=np.float32).T
    t2a = np.asarray(text_to_artifact)
    n = sim.shape[0]
    pos = np.empty(n, dtype=np.int64)
    for i in range(n):
        ranked = t2a[np.argsort(-sim[i])]
        pos[i] = np.where(ranked == i)[0][0] + 1
    out = {f"T2T_R@{k}": float(np.sum(pos <= k)) / n * 100 for k in k_values}
    out["T2T_MRR"] = np.mean(1.0 / pos) * 100
    out["T2T_Mean_Rank"] = np.mean(pos)
    return out


def canon_grouped_rank(scores: np.ndarray, text_to_artifact, query_artifact=None) -> np.ndarray:
    """Canonical form: 1 + #candidates ahead of the query's best positive under (score desc, index asc)."""
    t2a = np.asarray(text_to_artifact)
    n = scores.shape[0]
    qa = np.arange(n) if query_artifact is None else np.asarray(query_artifact)
    order = np.argsort(-scores, axis=1, kind="stable")
    return np.array([int(np.where(t2a[order[i]] == qa[i])[0][0]) + 1 for i in range(n)], dtype=np.int64)
