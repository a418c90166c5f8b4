"""Inference-side drop-in for the reference's learned fusion heads (`src/clip/model/fusion_model.py`) and their
evaluation loop (`src/clip/eval/evaluator_fusion.py:76-132`).

The reference scores 50 queries x 500 candidates per block on the GPU, copies every block into an (N, N) fp32 host
matrix and then sorts its rows twice.  The gated heads are `scores = gate(q)*T2I + (1 - gate(q))*T2T`: a PER-QUERY
weight pair, which here rides in the epilogue of the same fused scan (one TMEM lane is one query) -- no blocks, no
host matrix, no sort.  Only inference is in scope: parameters are passed in as arrays (e.g. from a checkpoint's
state_dict); training the heads stays with the reference.

* `SimpleGatedFusion` / `SimpleGatedFusionWithBias` (`fusion_model.py:182-196`, `:9-23`): gate = sigmoid(q.w + b),
  computed by `kemr_gate_linear` in fp32.
* `GatedFusionHead` (`:136-180`): the gate is a 2-layer MLP of the query; a (N,768)x(768,128) product is a plain
  library GEMM, so it runs through torch, and its output enters as per-query weights.
* `BilinearFusionHead` (`:198-240`): the galleries are projected once (`W_image`, `W_target`, plain GEMMs) and the
  result is a scalar-weighted fused scan with `sigmoid(alpha)`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine
from .metrics import _metrics_from_ranks


class _GatedBase:
    """scores(q, j) = gate(q) * <q, image_j> + (1 - gate(q)) * <q, target_j>."""

    def gate_weights(self, q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def search(self, query_embeddings, image_embeddings, target_embeddings, k: int = 10):
        """Top-k candidates per query under the head's score: (idx int64 [N,k], score f64 [N,k]) on the device."""
        q, img, tgt = (engine.quantize(x) for x in (query_embeddings, image_embeddings, target_embeddings))
        wa, wb = self.gate_weights(q)
        return engine.scan_topk(q, img, tgt, wa, wb, k=k)

    def evaluate(self, query_embeddings, image_embeddings, target_embeddings,
                 k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
        """Metrics the reference gets from `compute_retrieval_metrics_fusion(similarity_matrix)` after its block
        loop (`evaluator_fusion.py:126-132`): target of query i is candidate i."""
        q, img, tgt = (engine.quantize(x) for x in (query_embeddings, image_embeddings, target_embeddings))
        wa, wb = self.gate_weights(q)
        tidx = torch.arange(q.shape[0], device=q.device, dtype=torch.int64)
        ranks = engine.rank_targets(q, img, tgt, tidx, wa, wb)
        return _metrics_from_ranks(ranks, k_values, True, True)


class SimpleGatedFusion(_GatedBase):
    """Reference `fusion_model.py:182-196` (defaults: weight = ones, bias = 0)."""

    def __init__(self, query_weight: Optional[Sequence[float]] = None, bias: float = 0.0, embed_dim: int = 768):
        self.query_weight = np.ones(embed_dim, np.float32) if query_weight is None else np.asarray(query_weight, np.float32)
        self.bias = float(bias)

    def gate_weights(self, q):
        return engine.gate_linear(q, self.query_weight, self.bias)


class SimpleGatedFusionWithBias(SimpleGatedFusion):
    """Reference `fusion_model.py:9-23` (defaults: weight = zeros, bias = -2, i.e. gate ~ 0.12)."""

    def __init__(self, query_weight: Optional[Sequence[float]] = None, bias: float = -2.0, embed_dim: int = 768):
        super().__init__(np.zeros(embed_dim, np.float32) if query_weight is None else query_weight, bias, embed_dim)


class GatedFusionHead(_GatedBase):
    """Reference `fusion_model.py:136-180`: gate = sigmoid(W2 . relu(W1 q + b1) + b2) (dropout is inactive at
    inference).  w1 [128, D], b1 [128], w2 [1, 128] or [128], b2 scalar -- torch `nn.Linear` layouts."""

    def __init__(self, w1, b1, w2, b2):
        self.w1 = torch.as_tensor(np.asarray(w1, np.float32))
        self.b1 = torch.as_tensor(np.asarray(b1, np.float32))
        self.w2 = torch.as_tensor(np.asarray(w2, np.float32)).reshape(-1)
        self.b2 = float(np.asarray(b2).reshape(-1)[0])

    def gate_weights(self, q):
        dev = q.device
        h = torch.relu(q.float() @ self.w1.to(dev).T + self.b1.to(dev))
        gate = torch.sigmoid(h @ self.w2.to(dev) + self.b2)                     # fp32, like the reference
        return gate.double(), (1.0 - gate).double()


class BilinearFusionHead:
    """Reference `fusion_model.py:198-240`: alpha' * q.(W_i img) + (1 - alpha') * q.(W_t tgt), alpha' = sigmoid(alpha).
    The projected galleries are computed once and stay resident (bf16, like every gallery of the engine)."""

    def __init__(self, w_image, w_target, alpha: float = 0.5):
        self.w_image = torch.as_tensor(np.asarray(w_image, np.float32))
        self.w_target = torch.as_tensor(np.asarray(w_target, np.float32))
        a = 1.0 / (1.0 + np.exp(-np.float32(alpha)))
        self.w_i, self.w_t = float(np.float32(a)), float(np.float32(1.0) - np.float32(a))
        self._proj = None

    def project(self, image_embeddings, target_embeddings):
        img = torch.as_tensor(np.asarray(image_embeddings, np.float32)).cuda()
        tgt = torch.as_tensor(np.asarray(target_embeddings, np.float32)).cuda()
        self._proj = (engine.quantize(img @ self.w_image.cuda().T), engine.quantize(tgt @ self.w_target.cuda().T))
        return self._proj

    def search(self, query_embeddings, k: int = 10):
        img, tgt = self._proj
        return engine.scan_topk(engine.quantize(query_embeddings), img, tgt, self.w_i, self.w_t, k=k)

    def evaluate(self, query_embeddings, k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
        img, tgt = self._proj
        q = engine.quantize(query_embeddings)
        tidx = torch.arange(q.shape[0], device=q.device, dtype=torch.int64)
        return _metrics_from_ranks(engine.rank_targets(q, img, tgt, tidx, self.w_i, self.w_t), k_values, True, True)
