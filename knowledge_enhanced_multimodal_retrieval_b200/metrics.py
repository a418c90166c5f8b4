"""Drop-in for the reference's `src/clip/eval/metrics.py` (same names, arguments, defaults,
dict keys and units), computed on the B200 through libkemr.so.

Where the reference materialises an (N, M) fp32 similarity matrix and fully sorts every row
twice (`metrics.py:34,62`), this module never builds the matrix for the embedding-taking
entry points: the target's score is computed first, then one fused scan counts the rows that
outrank it (`engine.rank_targets`).  Recall@K / MRR / Mean_Rank follow from the ranks through
a fused device reduction that reproduces numpy's summation order, so the returned
`np.float64` values are bit-identical to the reference's whenever the ranks are.
"""
from __future__ import annotations

import logging
from typing import Dict, List

import numpy as np
import torch

from . import engine

logger = logging.getLogger(__name__)


def _metrics_from_ranks(ranks: torch.Tensor, k_values, compute_recall=True, compute_mrr=True,
                        prefix: str = "") -> Dict[str, float]:
    """`metrics.py:41-42,70-71` on 1-based ranks (device reduction, numpy-ordered sums)."""
    n = np.float64(ranks.numel())
    ks = list(k_values) if compute_recall else []
    hits, rank_sum, rr_sum = engine.metrics_reduce(ranks, ks)
    out: Dict[str, float] = {}
    for k, h in zip(ks, hits):
        out[f"R@{k}"] = np.float64(h) / n * 100.0
    if compute_mrr:
        out["MRR"] = np.float64(rr_sum) / n * 100.0
        out["Mean_Rank"] = np.float64(rank_sum) / n
    return {(f"{prefix}_{k}" if prefix else k): v for k, v in out.items()}


def compute_recall_at_k(similarity_matrix, k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
    """Recall@K of a caller-supplied similarity matrix (reference `metrics.py:13-44`)."""
    ranks = engine.matrix_rank(similarity_matrix)
    return _metrics_from_ranks(ranks, k_values, True, False)


def compute_mrr_and_mean_rank(similarity_matrix) -> Dict[str, float]:
    """MRR (percent) and Mean_Rank of a caller-supplied matrix (reference `metrics.py:47-76`)."""
    ranks = engine.matrix_rank(similarity_matrix)
    return _metrics_from_ranks(ranks, [], False, True)


def _rank_metrics(q, gal_a, gal_b, w_a, w_b, prefix, k_values, compute_recall, compute_mrr):
    n, m = q.shape[0], gal_a.shape[0]
    if not (compute_recall or compute_mrr):
        return {}
    nv = min(n, m)
    tidx = torch.arange(nv, device=q.device, dtype=torch.int64)         # metrics.py:37
    ranks = engine.rank_targets(q[:nv].contiguous() if nv < n else q, gal_a, gal_b, tidx, w_a, w_b)
    if nv < n:
        # more queries than candidates: column i does not exist for i >= M, so the reference finds no match --
        # never a recall hit, position argmax(all False) + 1 = 1 (metrics.py:41,68); rank 0 encodes exactly that
        ranks = torch.cat([ranks, torch.zeros(n - nv, dtype=ranks.dtype, device=ranks.device)])
    return _metrics_from_ranks(ranks, k_values, compute_recall, compute_mrr, prefix)


def compute_retrieval_metrics(query_embeddings, candidate_embeddings, prefix: str = "",
                              k_values: List[int] = [1, 5, 10, 20], compute_recall: bool = True,
                              compute_mrr: bool = True) -> Dict[str, float]:
    """Reference `metrics.py:79-116`; the sgemm at :102 and both argsorts are one fused scan."""
    q = engine.quantize(query_embeddings)
    c = engine.quantize(candidate_embeddings)
    return _rank_metrics(q, c, None, 1.0, 0.0, prefix, k_values, compute_recall, compute_mrr)


def compute_retrieval_metrics_final(query_embeddings, target_embeddings, image_embeddings, prefix: str = "",
                                    k_values: List[int] = [1, 5, 10, 20], compute_recall: bool = True,
                                    compute_mrr: bool = True, t2i_weight: float = 0.5,
                                    t2t_weight: float = 0.5) -> Dict[str, float]:
    """Reference `metrics.py:119-162`: w_i*(Q.I^T) + w_t*(Q.T^T) fused in the scan epilogue."""
    print("compute_retrieval_metrics_final:", t2i_weight, t2t_weight)     # metrics.py:147
    q = engine.quantize(query_embeddings)
    img = engine.quantize(image_embeddings)
    tgt = engine.quantize(target_embeddings)
    return _rank_metrics(q, img, tgt, t2i_weight, t2t_weight, prefix, k_values, compute_recall, compute_mrr)


def compute_retrieval_metrics_fusion(similarity_matrix, prefix: str = "", k_values: List[int] = [1, 5, 10, 20],
                                     compute_recall: bool = True, compute_mrr: bool = True) -> Dict[str, float]:
    """Reference `metrics.py:165-185`."""
    if not (compute_recall or compute_mrr):
        return {}
    ranks = engine.matrix_rank(similarity_matrix)
    return _metrics_from_ranks(ranks, k_values, compute_recall, compute_mrr, prefix)


def compute_all_retrieval_metrics(query_embeddings, target_embeddings, image_embeddings,
                                  k_values: List[int] = [1, 5, 10, 20],
                                  tasks: List[str] = ["T2I", "I2T", "T2T"], compute_recall: bool = True,
                                  compute_mrr: bool = True) -> Dict[str, float]:
    """Reference `metrics.py:188-252`: T2I q->img, I2T img->tgt, T2T q->tgt.
    Each embedding set is quantised and uploaded once and stays resident for all tasks."""
    q = engine.quantize(query_embeddings)
    tgt = engine.quantize(target_embeddings)
    img = engine.quantize(image_embeddings)
    pairs = {"T2I": (q, img), "I2T": (img, tgt), "T2T": (q, tgt)}
    metrics: Dict[str, float] = {}
    for name in ("T2I", "I2T", "T2T"):
        if name in tasks:
            a, b = pairs[name]
            metrics.update(_rank_metrics(a, b, None, 1.0, 0.0, name, k_values, compute_recall, compute_mrr))
    return metrics


def compute_training_metrics(query_embeddings, target_embeddings, image_embeddings,
                             tasks: List[str] = ["T2I", "I2T", "T2T"]) -> Dict[str, float]:
    """Reference `metrics.py:256-282`: MRR / Mean_Rank only (early stopping)."""
    return compute_all_retrieval_metrics(query_embeddings, target_embeddings, image_embeddings,
                                         tasks=tasks, compute_recall=False, compute_mrr=True)


def grouped_ranks(q: torch.Tensor, cand: torch.Tensor, candidate_to_artifact, query_artifact=None) -> torch.Tensor:
    """1-based position of the first candidate belonging to the query's artefact (many candidates per artefact:
    the N x 4 pools of `baselines/evaluate_text_models.py:176-186,237-249`).  The best positive of each query is
    found on canonical scores of its (few) positives, then one fused scan counts the candidates ranked ahead of it."""
    c2a = np.asarray(candidate_to_artifact, dtype=np.int64)
    n = q.shape[0]
    qa = np.arange(n, dtype=np.int64) if query_artifact is None else np.asarray(query_artifact, dtype=np.int64)
    order = np.argsort(c2a, kind="stable")
    first = np.searchsorted(c2a[order], qa, side="left")
    last = np.searchsorted(c2a[order], qa, side="right")
    sizes = last - first
    if (sizes <= 0).any():
        raise AssertionError("every query needs at least one candidate of its artefact")
    g = int(sizes.max())
    pos = np.full((n, g), -1, dtype=np.int64)                       # candidate rows of each query's artefact
    for j in range(g):
        m = sizes > j
        pos[m, j] = order[first[m] + j]
    pos_d = torch.from_numpy(pos).to(q.device)
    valid = pos_d >= 0
    pq = torch.arange(n, device=q.device, dtype=torch.int32).repeat_interleave(g)
    sc = engine.score_pairs(q, cand, None, pq, pos_d.clamp(min=0).flatten()).view(n, g)
    sc = torch.where(valid, sc, torch.full_like(sc, float("-inf")))
    best = sc.max(dim=1, keepdim=True).values
    big = torch.iinfo(torch.int64).max
    tidx = torch.where(valid & (sc == best), pos_d, torch.full_like(pos_d, big)).min(dim=1).values   # ties: lowest row
    t = sc.gather(1, (pos_d == tidx[:, None]).to(torch.int64).argmax(dim=1, keepdim=True)).flatten().contiguous()
    return engine.rank_count(q, cand, None, t, tidx.contiguous()) + 1


def compute_grouped_retrieval_metrics(query_embeddings, candidate_embeddings, candidate_to_artifact,
                                      query_artifact=None, prefix: str = "T2T",
                                      k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
    """Recall@K / MRR / Mean_Rank with several correct candidates per query -- the `single` mode of
    `baselines/evaluate_text_models.py:171-224` (`{prefix}_R@k`, `{prefix}_MRR`, `{prefix}_Mean_Rank`).  For its
    `multi` mode (:226-281) pass the five query variants stacked and `query_artifact = tile(arange(N), 5)` against
    each pool in turn, or call `grouped_ranks` per variant and reduce the concatenated ranks once."""
    q = engine.quantize(query_embeddings)
    c = engine.quantize(candidate_embeddings)
    ranks = grouped_ranks(q, c, candidate_to_artifact, query_artifact)
    return _metrics_from_ranks(ranks, k_values, True, True, prefix)


def _deprecated(name):
    logger.warning("%s is DEPRECATED. Use compute_all_retrieval_metrics / compute_training_metrics "
                   "with separate query and target embeddings.", name)


def compute_metrics_multi_mode(image_embeddings, text_embeddings_by_variant) -> Dict[str, float]:
    """Reference `metrics.py:285-304` (self-retrieval shim)."""
    _deprecated("compute_metrics_multi_mode")
    t = text_embeddings_by_variant[0]
    return compute_all_retrieval_metrics(t, t, image_embeddings, tasks=["T2I", "I2T", "T2T"])


def compute_metrics_single_4train(image_embeddings, text_embeddings_by_variant) -> Dict[str, float]:
    """Reference `metrics.py:307-328`."""
    _deprecated("compute_metrics_single_4train")
    t = text_embeddings_by_variant[0]
    return compute_training_metrics(t, t, image_embeddings, tasks=["T2I", "I2T", "T2T"])


def compute_metrics_multi_4train(image_embeddings, text_embeddings_by_variant) -> Dict[str, float]:
    """Reference `metrics.py:331-352`."""
    _deprecated("compute_metrics_multi_4train")
    t = text_embeddings_by_variant[0]
    return compute_training_metrics(t, t, image_embeddings, tasks=["T2I", "I2T", "T2T"])
