"""Deterministic synthetic embeddings and knowledge-graph hit sets.

Everything here is host-side numpy so the same seeded values reach the oracle, the
golden-fixture generator and the GPU engine.  Embeddings are L2-normalised in fp32
and then rounded to bf16 (round-to-nearest-even); the oracle consumes the very same
values upcast to fp32, so no quantisation error separates the two sides.

Shapes follow the reference's evaluation loop (`src/clip/eval/evaluator.py:140-143`):
three `(N, D)` arrays -- query, target-text and image embeddings -- where row ``i`` of
each belongs to artefact ``i`` (the ground truth used at `metrics.py:37`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


# --------------------------------------------------------------------------- bf16 helpers
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round-to-nearest-even, NaN kept quiet."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    rounding = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    out = ((u + rounding) >> np.uint32(16)).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out[nan] = ((u[nan] >> np.uint32(16)) | np.uint32(0x0040)).astype(np.uint16)
    return out


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    """bf16 bit patterns (uint16) -> exact fp32 values."""
    b = np.ascontiguousarray(b, dtype=np.uint16)
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32)


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 values rounded to the nearest bf16-representable fp32 value."""
    return bf16_bits_to_f32(f32_to_bf16_bits(x))


def l2_normalize(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    n = np.sqrt((x.astype(np.float32) ** 2).sum(axis=1, keepdims=True, dtype=np.float32))
    return (x / n).astype(np.float32)


# --------------------------------------------------------------------------- datasets
@dataclass
class SyntheticRetrievalSet:
    """One synthetic evaluation set; all arrays hold bf16-representable fp32 values."""

    query: np.ndarray            # (Q, D) fp32
    image: np.ndarray            # (M, D) fp32   T2I gallery
    target: Optional[np.ndarray]  # (M, D) fp32   T2T gallery (None for single-gallery sets)
    target_idx: np.ndarray       # (Q,) int64    ground-truth gallery row of each query
    uuids: List[str] = field(default_factory=list)          # gallery uuids, len M
    query_uuids: List[str] = field(default_factory=list)    # len Q
    kg_results: Dict[str, List[str]] = field(default_factory=dict)  # query uuid -> artefact URIs

    @property
    def Q(self) -> int:
        return self.query.shape[0]

    @property
    def M(self) -> int:
        return self.image.shape[0]

    @property
    def D(self) -> int:
        return self.image.shape[1]


def make_gallery(M: int, D: int, seed: int) -> np.ndarray:
    """Gallery rows x ~ N(0, I_D) -> x/||x|| (fp32) -> bf16-rounded fp32."""
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((M, D), dtype=np.float32)
    return round_to_bf16(l2_normalize(g))


def make_queries(galleries: Tuple[np.ndarray, ...], target_idx: np.ndarray, lam: float,
                 seed: int) -> np.ndarray:
    """queries = normalize(lam * mean(gallery[target]) + eps), eps ~ N(0, I/D)."""
    D = galleries[0].shape[1]
    rng = np.random.default_rng(seed)
    eps = rng.standard_normal((len(target_idx), D), dtype=np.float32) / np.float32(np.sqrt(D))
    base = sum(g[target_idx] for g in galleries) / np.float32(len(galleries))
    return round_to_bf16(l2_normalize(np.float32(lam) * base + eps))


def make_kg_results(Q: int, M: int, target_idx: np.ndarray, seed: int, mean_hits: float = 20.0,
                    p_target: float = 0.5, p_unknown: float = 0.02, p_dup: float = 0.02
                    ) -> Tuple[Dict[str, List[str]], List[str], List[str]]:
    """Per query a Poisson(mean_hits) subset of gallery uuids (+ the true target w.p. 0.5).

    Half of the entries are wrapped as full URIs (exercises the ``split('/')[-1]`` at
    `src/clip/eval/fusion.py:76`), a few are unknown uuids (ignored, `fusion.py:78`) and a
    few are duplicated (idempotent in `weighted_fusion`, additive twice in
    `additive_bonus_fusion`, `fusion.py:130`).  Some queries have no entry at all.
    """
    rng = np.random.default_rng(seed)
    uuids = [f"u{j:07d}" for j in range(M)]
    query_uuids = [f"q{i:07d}" for i in range(Q)]
    results: Dict[str, List[str]] = {}
    for i in range(Q):
        if rng.random() < 0.1:
            continue                       # query absent from the results -> no boost
        n = int(rng.poisson(mean_hits))
        hits = list(rng.integers(0, M, size=n))
        if rng.random() < p_target:
            hits.append(int(target_idx[i]))
        out: List[str] = []
        for h in hits:
            u = uuids[int(h)]
            if rng.random() < p_unknown:
                u = f"zz{int(h)}"           # not in the gallery
            out.append(f"http://kg.example/artefact/{u}" if rng.random() < 0.5 else u)
            if rng.random() < p_dup:
                out.append(out[-1])
        rng.shuffle(out)
        results[query_uuids[i]] = out
    return results, query_uuids, uuids


def make_retrieval_set(Q: int, M: int, D: int, seed: int, fused: bool = True, lam: float = 0.1,
                       with_kg: bool = False, diagonal: bool = True) -> SyntheticRetrievalSet:
    """Synthetic set in the shape of the reference's eval arrays.

    ``diagonal=True`` makes query ``i`` target gallery row ``i`` (the only ground-truth
    convention the reference's metrics support, `metrics.py:37`); it requires Q <= M.
    """
    image = make_gallery(M, D, seed)
    target = make_gallery(M, D, seed + 1) if fused else None
    if diagonal:
        assert Q <= M
        tidx = np.arange(Q, dtype=np.int64)
    else:
        tidx = np.random.default_rng(seed + 2).integers(0, M, size=Q).astype(np.int64)
    gal = (image, target) if fused else (image,)
    query = make_queries(gal, tidx, lam, seed + 3)
    s = SyntheticRetrievalSet(query=query, image=image, target=target, target_idx=tidx)
    if with_kg:
        s.kg_results, s.query_uuids, s.uuids = make_kg_results(Q, M, tidx, seed + 4)
    else:
        s.uuids = [f"u{j:07d}" for j in range(M)]
        s.query_uuids = [f"q{i:07d}" for i in range(Q)]
    return s


# Named configurations of BASELINE.json (C4/C5 galleries are generated on the device).
CONFIGS = {
    "c1": dict(Q=4300, M=43000, D=512, seed=0, fused=False),
    "c2": dict(Q=1000, M=43000, D=768, seed=1, fused=True),
    "c3": dict(Q=1, M=43000, D=768, seed=1, fused=True, with_kg=True),
}
