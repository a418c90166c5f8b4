"""Drop-in for the reference's serving engine `src/retrieval.py` and its CLIP shim
`src/clip/clip_retrieval.py`.

`RetrievalEngine` keeps the reference's method names, arguments, defaults and return shapes.
The two network-bound collaborators are injectable (the reference builds them in `__init__`
from environment variables, `retrieval.py:13-21`): `t2s_retriever` is any object with
`.retrieval(query) -> List[uuid]` (the Text2SPARQL pipeline is out of scope here), and
`clip_retriever` is a `CLIPRetrieval` whose `.retriever` is the B200-resident `CLIPRetriever`
below instead of code downloaded from the HF hub.

`CLIPRetriever.search` is PARITY-UNPINNED: the reference's implementation is not in its tree
(`clip_retrieval.py:15-23`).  It is defined here as the fused T2I/T2T scan the evaluation code
uses (`metrics.py:145-148`): score = alpha*T2I + (1-alpha)*T2T, top-`top_k`, descending.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np

from .index import GalleryIndex


class CLIPRetriever:
    """Resident gallery + text encoder.  `encode_text(str) -> (D,) or (1, D)` fp32, L2-normalised."""

    def __init__(self, index: GalleryIndex, encode_text: Optional[Callable] = None, top_k: int = 100):
        if index.uuids is None:
            raise ValueError("CLIPRetriever needs a GalleryIndex with uuids")
        self.index = index
        self.encode_text = encode_text
        self.top_k = top_k

    def search(self, query, alpha: float = 0.5, top_k: Optional[int] = None) -> List[Dict]:
        k = min(top_k or self.top_k, self.index.M)
        if isinstance(query, str) and self.encode_text is None:
            raise ValueError("no text encoder configured; pass a query embedding")
        emb = self.encode_text(query) if isinstance(query, str) else query
        emb = np.asarray(emb, dtype=np.float32).reshape(1, -1)
        if self.index.target is not None:
            idx, score = self.index.search(emb, k=k, t2i_weight=alpha, t2t_weight=1.0 - alpha)
        else:
            idx, score = self.index.search(emb, k=k)
        idx = idx[0].cpu().numpy()
        score = score[0].cpu().numpy()
        # a shard of a row-partitioned gallery returns GLOBAL ids (idx_base + row) but holds its own uuid slice
        base = self.index.idx_base
        return [{"uuid": self.index.uuids[int(j) - base], "score": float(s)} for j, s in zip(idx, score) if j >= 0]


class CLIPRetrieval:
    """Reference `src/clip/clip_retrieval.py:10-40`: `.retrieval(query, alpha)` -> ranked list."""

    def __init__(self, model_name=None, retriever: Optional[CLIPRetriever] = None):
        if retriever is None:
            raise ValueError("pass retriever=CLIPRetriever(...): the engine does not download code or embeddings")
        self.model_name = model_name
        self.retriever = retriever

    def retrieval(self, query, alpha: float = 0.5):
        return self.retriever.search(query, alpha=alpha)


class RetrievalEngine:

    def __init__(self, clip_retriever=None, t2s_retriever=None):
        self.clip_retriever = clip_retriever
        self.t2s_retriever = t2s_retriever

    def _fuse_clip_sparql_linear(self, clip_results: List[Dict], sparql_results: List[str],
                                 alpha: float = 0.8, beta: float = 0.2) -> List[Dict]:
        """Reference `retrieval.py:23-76`: python-float alpha*clip + beta*[uuid in sparql], rounded to
        4 decimals, stable descending sort (ties keep CLIP order); SPARQL-only uuids never appear."""
        if not clip_results:
            return []
        hit = set(sparql_results)
        fused = [{"uuid": r["uuid"],
                  "score": round(alpha * r["score"] + beta * (1.0 if r["uuid"] in hit else 0.0), 4)}
                 for r in clip_results]
        fused.sort(key=lambda r: r["score"], reverse=True)
        return fused

    def retrieve_text(self, query, alpha: float = 0.8, beta: float = 0.2, alpha_clip: float = 0.5,
                      threshold: float = 0):
        """Reference `retrieval.py:79-95`."""
        clip_results = self.clip_retriever.retrieval(query, alpha=alpha_clip)
        t2s_results = self.t2s_retriever.retrieval(query)
        fused = self._fuse_clip_sparql_linear(clip_results=clip_results, sparql_results=t2s_results,
                                              alpha=alpha, beta=beta)
        return [{"uuid": r["uuid"], "score": r["score"]} for r in fused if r.get("score", 0) >= threshold]

    def retrieve_text_noknowledge(self, query, alpha: float = 0.8, beta: float = 0.2, alpha_clip: float = 0.5,
                                  threshold: float = 0):
        """Reference `retrieval.py:97-107`: CLIP scores pass through unscaled and unrounded."""
        results = self.clip_retriever.retrieval(query, alpha=alpha_clip)
        return [{"uuid": r["uuid"], "score": r["score"]} for r in results if r.get("score", 0) >= threshold]
