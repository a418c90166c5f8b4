"""Drop-in for the reference's `src/clip/eval/fusion.py`, computed on the B200.

Matrix-taking functions keep the reference's contract exactly -- a NEW fp32 (N, M) array,
input untouched, same assertions and ValueError -- and are bit-identical to numpy's fp32
arithmetic (kemr_matrix_fuse).  `evaluate_fused` is the matrix-free route for callers that
only want the metrics of the fused ranking (the alpha sweep of `evaluator.py:164-218`): the
KG boost rides the scan as a sparse side path and no (N, M) array is ever built.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import engine
from .metrics import _metrics_from_ranks, compute_mrr_and_mean_rank, compute_recall_at_k  # noqa: F401  (re-exported like the reference's fusion.py:3)

_DEFAULT_OMEGA = {1: 1.0, 5: 0.8, 20: 0.5, 50: 0.3, float("inf"): 0.1}      # fusion.py:164-170


def evaluate_retrieval(similarity_matrix) -> Dict[str, float]:
    """Reference `fusion.py:6-20`: R@{1,5,10,20} + MRR + Mean_Rank of a matrix, printed."""
    ranks = engine.matrix_rank(similarity_matrix)
    metrics = _metrics_from_ranks(ranks, [1, 5, 10, 20], True, True)
    print("evaluate_retrieval:", metrics)
    return metrics


def _check_shapes(S, query_uuids, artefact_uuids):
    assert S.shape[0] == len(query_uuids), \
        f"Similarity matrix rows ({S.shape[0]}) != query_uuids length ({len(query_uuids)})"
    assert S.shape[1] == len(artefact_uuids), \
        f"Similarity matrix cols ({S.shape[1]}) != artefact_uuids length ({len(artefact_uuids)})"


def _omega(size: int, thresholds) -> float:
    for thr, w in sorted(thresholds.items()):          # fusion.py:191-196
        if size <= thr:
            return w
    return 0.0


def _csr(cols_per_query, add_per_query):
    rowptr = np.zeros(len(cols_per_query) + 1, dtype=np.int64)
    for i, c in enumerate(cols_per_query):
        rowptr[i + 1] = rowptr[i] + len(c)
    col = np.fromiter((c for cs in cols_per_query for c in cs), dtype=np.int32, count=int(rowptr[-1]))
    add = np.fromiter((a for as_ in add_per_query for a in as_), dtype=np.float32, count=int(rowptr[-1]))
    return rowptr, col, add


def _finish(out: torch.Tensor, like):
    return out if isinstance(like, torch.Tensor) else out.cpu().numpy()


def weighted_fusion(clip_similarity_matrix, text2sparql_results: Dict[str, List[str]], query_uuids: List[str],
                    artefact_uuids: List[str], alpha: float = 0.7, sparql_weight: float = 0.3):
    """alpha*S + w*I(d in R_SPARQL(q)) (reference `fusion.py:22-85`), dense fp32 result."""
    _check_shapes(clip_similarity_matrix, query_uuids, artefact_uuids)
    if not np.isclose(alpha + sparql_weight, 1.0):
        print(f"Warning: alpha ({alpha}) + sparql_weight ({sparql_weight}) != 1.0, normalizing...")
        total = alpha + sparql_weight
        alpha, sparql_weight = alpha / total, sparql_weight / total
    cols, _ = engine.kg_pairs(text2sparql_results, query_uuids, artefact_uuids)
    cols = [list(dict.fromkeys(c)) for c in cols]                      # indicator: duplicates idempotent
    w32 = np.float32(sparql_weight)
    rowptr, col, add = _csr(cols, [[w32] * len(c) for c in cols])
    out = engine.matrix_fuse(clip_similarity_matrix, True, np.float32(alpha), rowptr, col, add)
    return _finish(out, clip_similarity_matrix)


def additive_bonus_fusion(clip_similarity_matrix, text2sparql_results: Dict[str, List[str]],
                          query_uuids: List[str], artefact_uuids: List[str], delta: float = 0.5):
    """S + delta per listed hit, duplicates add again (reference `fusion.py:88-132`)."""
    _check_shapes(clip_similarity_matrix, query_uuids, artefact_uuids)
    cols, _ = engine.kg_pairs(text2sparql_results, query_uuids, artefact_uuids)
    d32 = np.float32(delta)
    rowptr, col, add = _csr(cols, [[d32] * len(c) for c in cols])
    out = engine.matrix_fuse(clip_similarity_matrix, False, 1.0, rowptr, col, add)
    return _finish(out, clip_similarity_matrix)


def adaptive_additive_fusion(clip_similarity_matrix, text2sparql_results: Dict[str, List[str]],
                             query_uuids: List[str], artefact_uuids: List[str], delta: float = 0.5,
                             size_thresholds: Dict[str, float] = None):
    """S + delta*omega(|R_SPARQL(q)|) per listed hit (reference `fusion.py:135-206`)."""
    if size_thresholds is None:
        size_thresholds = dict(_DEFAULT_OMEGA)
    _check_shapes(clip_similarity_matrix, query_uuids, artefact_uuids)
    cols, sizes = engine.kg_pairs(text2sparql_results, query_uuids, artefact_uuids)
    adds = [[np.float32(delta * _omega(sizes[i], size_thresholds))] * len(c) for i, c in enumerate(cols)]
    rowptr, col, add = _csr(cols, adds)
    out = engine.matrix_fuse(clip_similarity_matrix, False, 1.0, rowptr, col, add)
    return _finish(out, clip_similarity_matrix)


def fuse_clip_and_text2sparql(clip_similarity_matrix, text2sparql_results: Dict[str, List[str]],
                              query_uuids: List[str], artefact_uuids: List[str],
                              fusion_strategy: str = "weighted", fusion_params: Dict = None):
    """Strategy dispatch with the reference's defaults (reference `fusion.py:209-276`)."""
    if fusion_params is None:
        fusion_params = {}
    if fusion_strategy == "weighted":
        return weighted_fusion(clip_similarity_matrix, text2sparql_results, query_uuids, artefact_uuids,
                               alpha=fusion_params.get("alpha", 0.7),
                               sparql_weight=fusion_params.get("sparql_weight", 0.3))
    if fusion_strategy == "additive":
        return additive_bonus_fusion(clip_similarity_matrix, text2sparql_results, query_uuids, artefact_uuids,
                                     delta=fusion_params.get("delta", 0.5))
    if fusion_strategy == "adaptive":
        return adaptive_additive_fusion(clip_similarity_matrix, text2sparql_results, query_uuids,
                                        artefact_uuids, delta=fusion_params.get("delta", 0.5),
                                        size_thresholds=fusion_params.get("size_thresholds", None))
    raise ValueError(f"Unknown fusion strategy: {fusion_strategy}")


# --------------------------------------------------------------------------- matrix-free route
def kg_hits_for_strategy(text2sparql_results, query_uuids, artefact_uuids, fusion_strategy="weighted",
                         fusion_params: Optional[Dict] = None):
    """(alpha, KGHits) such that final = alpha*clip + bonus reproduces the chosen strategy."""
    p = fusion_params or {}
    if fusion_strategy == "weighted":
        alpha, w = p.get("alpha", 0.7), p.get("sparql_weight", 0.3)
        if not np.isclose(alpha + w, 1.0):
            tot = alpha + w
            alpha, w = alpha / tot, w / tot
        return alpha, engine.build_hits(text2sparql_results, query_uuids, artefact_uuids, lambda i, n: w, True)
    if fusion_strategy == "additive":
        d = p.get("delta", 0.5)
        return 1.0, engine.build_hits(text2sparql_results, query_uuids, artefact_uuids, lambda i, n: d, False)
    if fusion_strategy == "adaptive":
        d = p.get("delta", 0.5)
        th = p.get("size_thresholds") or dict(_DEFAULT_OMEGA)
        return 1.0, engine.build_hits(text2sparql_results, query_uuids, artefact_uuids,
                                      lambda i, n: d * _omega(n, th), False)
    raise ValueError(f"Unknown fusion strategy: {fusion_strategy}")


def evaluate_fused(query_embeddings, target_embeddings, image_embeddings, text2sparql_results,
                   query_uuids, artefact_uuids, t2i_weight: float = 0.5, t2t_weight: float = 0.5,
                   fusion_strategy: str = "weighted", fusion_params: Optional[Dict] = None,
                   k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
    """Metrics of `evaluate_retrieval(fuse_clip_and_text2sparql(w_i*QI^T + w_t*QT^T, ...))`
    (the loop body of `evaluator.py:176-190`) without materialising any (N, M) matrix."""
    q = engine.quantize(query_embeddings)
    img = engine.quantize(image_embeddings)
    tgt = engine.quantize(target_embeddings) if target_embeddings is not None else None
    assert q.shape[0] == len(query_uuids) and img.shape[0] == len(artefact_uuids)
    alpha, hits = kg_hits_for_strategy(text2sparql_results, query_uuids, artefact_uuids, fusion_strategy,
                                       fusion_params)
    tidx = torch.arange(q.shape[0], device=q.device, dtype=torch.int64)
    if tgt is None:
        ranks = engine.rank_targets(q, img, None, tidx, 1.0, 0.0, alpha, hits)
    else:
        ranks = engine.rank_targets(q, img, tgt, tidx, t2i_weight, t2t_weight, alpha, hits)
    return _metrics_from_ranks(ranks, k_values, True, True)


def evaluate_weight_and_alpha_sweep(query_embeddings, target_embeddings, image_embeddings, text2sparql_results,
                                    uuid_list, weight_settings=((0.5, 0.5), (0.1, 0.9)),
                                    alphas=(0.9, 0.8, 0.7, 0.6, 0.5, 0.4, 0.3, 0.2, 0.1),
                                    k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, Dict[str, float]]:
    """The evaluation block of the reference driver (`evaluator.py:164-218`) in one call: for every
    (t2i_weight, t2t_weight) setting the metrics of T2I alone, T2T alone, the fused similarity, and of
    `fuse_clip_and_text2sparql(fused, results, uuid_list, uuid_list, "weighted", {alpha, 1 - alpha})` for every alpha
    -- 2 x (3 + 9) = 24 evaluations there, each building (N, N) matrices and sorting them twice.  Here the embeddings
    are quantised and uploaded once, the KG result lists are mapped to gallery rows once and stay on the device, and
    every evaluation is one target-score pass + one fused counting scan; no (N, N) matrix exists.
    Keys: "w{t2i}_{t2t}/T2I", ".../T2T", ".../Fused", ".../alpha{a}"."""
    from . import store
    q = engine.quantize(query_embeddings)
    img = engine.quantize(image_embeddings)
    tgt = engine.quantize(target_embeddings)
    n = q.shape[0]
    assert n == len(uuid_list) and img.shape[0] == len(uuid_list)
    tidx = torch.arange(n, device=q.device, dtype=torch.int64)
    lists = store.HitLists(store.IdMap(uuid_list), text2sparql_results, uuid_list)
    out: Dict[str, Dict[str, float]] = {}
    single = {"T2I": _metrics_from_ranks(engine.rank_targets(q, img, None, tidx), k_values, True, True),
              "T2T": _metrics_from_ranks(engine.rank_targets(q, tgt, None, tidx), k_values, True, True)}
    for wi, wt in weight_settings:
        tag = f"w{wi}_{wt}"
        out[f"{tag}/T2I"], out[f"{tag}/T2T"] = single["T2I"], single["T2T"]     # do not depend on the weights
        out[f"{tag}/Fused"] = _metrics_from_ranks(engine.rank_targets(q, img, tgt, tidx, wi, wt), k_values, True, True)
        for a in alphas:
            alpha, hits = lists.for_strategy("weighted", {"alpha": a, "sparql_weight": 1 - a})
            ranks = engine.rank_targets(q, img, tgt, tidx, wi, wt, alpha, hits)
            out[f"{tag}/alpha{a}"] = _metrics_from_ranks(ranks, k_values, True, True)
    return out
