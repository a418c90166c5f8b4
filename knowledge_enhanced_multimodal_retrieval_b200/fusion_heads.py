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
* `LinearFusionHead` (`:25-48`): a 2 -> 128 -> 1 MLP of the pair (T2I score, T2T score) for every (query, candidate)
  -- not a weighted sum, so it does not ride the top-k epilogue: both fp32 score matrices come from the scan kernels'
  dense mode and one streaming kernel (`kemr_matrix_mlp2`) applies the MLP; ranks / top-k via the matrix entry points.
* `CrossAttentionFusionHead` (`:51-133`): per PAIR a 2-token attention, an output projection and a 768-256-64-1 MLP
  -- ~0.8 MFLOP per pair of plain dense layers; evaluated in blocks with library GEMMs (torch) into a device-resident
  score matrix, then ranked by the matrix entry points (the reference does the same 50 x 500 blocks into a host
  matrix, `evaluator_fusion.py:76-121`).
`FusionModel` (`:243-332`) dispatches on `fusion_type` exactly like the reference's wrapper.
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



class LinearFusionHead:
    """Reference `fusion_model.py:25-48`: scores = Linear(128, 1)(relu(Linear(2, 128)([t2i, t2t]))) per element (dropout
    inactive at inference).  w1 [128, 2], b1 [128], w2 [1, 128] or [128], b2 scalar -- torch `nn.Linear` layouts."""

    def __init__(self, w1, b1, w2, b2):
        self.w1 = np.ascontiguousarray(np.asarray(w1, np.float32).reshape(-1, 2))
        self.b1 = np.ascontiguousarray(np.asarray(b1, np.float32).reshape(-1))
        self.w2 = np.ascontiguousarray(np.asarray(w2, np.float32).reshape(-1))
        self.b2 = float(np.asarray(b2).reshape(-1)[0])
        if not (len(self.w1) == len(self.b1) == len(self.w2)):
            raise engine.KemrError("LinearFusionHead: inconsistent parameter shapes")

    def fuse(self, t2i_sim: torch.Tensor, t2t_sim: torch.Tensor) -> torch.Tensor:
        """`forward(t2i_sim, t2t_sim)` of the reference on two fp32 CUDA matrices [N, M]."""
        a = t2i_sim.to(device="cuda", dtype=torch.float32).contiguous()
        b = t2t_sim.to(device="cuda", dtype=torch.float32).contiguous()
        if a.shape != b.shape or a.dim() != 2:
            raise engine.KemrError("LinearFusionHead: the two score matrices must have the same 2-D shape")
        out = torch.empty_like(a)
        dev = a.device
        w1, b1, w2 = (torch.from_numpy(x).to(dev) for x in (self.w1, self.b1, self.w2))
        engine._lib.check(engine._lib.load().kemr_matrix_mlp2(engine._ptr(a), engine._ptr(b), engine._ptr(out), a.shape[0], a.shape[1],
                                                            engine._ptr(w1), engine._ptr(b1), engine._ptr(w2), self.b2, len(self.b1),
                                                            engine._stream()))
        return out

    def scores(self, query_embeddings, image_embeddings, target_embeddings) -> torch.Tensor:
        """`FusionModel.forward` for fusion_type "linear" (`fusion_model.py:318-322`): both similarity matrices from the
        scan kernels' dense mode (fp32 accumulation), then the MLP."""
        q, img, tgt = (engine.quantize(x) for x in (query_embeddings, image_embeddings, target_embeddings))
        return self.fuse(engine.score_matrix(q, img), engine.score_matrix(q, tgt))

    def evaluate(self, query_embeddings, image_embeddings, target_embeddings,
                 k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
        ranks = engine.matrix_rank(self.scores(query_embeddings, image_embeddings, target_embeddings))
        return _metrics_from_ranks(ranks, k_values, True, True)

    def search(self, query_embeddings, image_embeddings, target_embeddings, k: int = 10):
        idx, val = engine.matrix_topk(self.scores(query_embeddings, image_embeddings, target_embeddings), k)
        return idx, val


class CrossAttentionFusionHead:
    """Reference `fusion_model.py:51-133`, inference only.  Parameters are the module's `state_dict()` arrays
    (`query_proj.*`, `image_proj.*`, `target_proj.*`, `cross_attn.in_proj_weight/bias`, `cross_attn.out_proj.*`,
    `score_mlp.{0,3,6}.*`).  Every layer is a plain dense layer, so the pair block runs on library GEMMs (torch); the
    galleries' key / value projections are computed once, the (N, M) scores stay on the device."""

    def __init__(self, state: Dict[str, np.ndarray], num_heads: int = 8):
        self.p = {k: torch.as_tensor(np.asarray(v, np.float32)) for k, v in state.items()}
        self.h = num_heads

    def scores(self, query_embeddings, image_embeddings, target_embeddings, block: int = 64) -> torch.Tensor:
        dev = torch.device("cuda")
        P = {k: v.to(dev) for k, v in self.p.items()}
        lin = lambda x, n: x @ P[n + ".weight"].T + P[n + ".bias"]                                  # noqa: E731
        q = lin(torch.as_tensor(np.asarray(query_embeddings, np.float32)).to(dev), "query_proj")
        img = lin(torch.as_tensor(np.asarray(image_embeddings, np.float32)).to(dev), "image_proj")
        tgt = lin(torch.as_tensor(np.asarray(target_embeddings, np.float32)).to(dev), "target_proj")
        D, H = q.shape[1], self.h
        hd = D // H
        Wi, bi = P["cross_attn.in_proj_weight"], P["cross_attn.in_proj_bias"]
        qh = (q @ Wi[:D].T + bi[:D]).view(-1, H, hd)                                              # (N, H, hd)
        kv = torch.stack([img, tgt], dim=1)                                                       # (M, 2, D)
        kh = (kv @ Wi[D:2 * D].T + bi[D:2 * D]).view(-1, 2, H, hd)                                # (M, 2, H, hd)
        vh = (kv @ Wi[2 * D:].T + bi[2 * D:]).view(-1, 2, H, hd)
        Wo, bo = P["cross_attn.out_proj.weight"], P["cross_attn.out_proj.bias"]
        # the output projection is linear: project each candidate's two value tokens per head once
        uo = torch.einsum("mthd,ehd->mthe", vh, Wo.view(D, H, hd))                                # (M, 2, H, D)
        N, M = q.shape[0], img.shape[0]
        out = torch.empty((N, M), dtype=torch.float32, device=dev)
        scale = 1.0 / float(hd) ** 0.5
        for n0 in range(0, N, block):
            qb = qh[n0:n0 + block]                                                                # (b, H, hd)
            logits = torch.einsum("bhd,mthd->bmht", qb, kh) * scale                               # (b, M, H, 2)
            att = torch.softmax(logits, dim=-1)
            a = torch.einsum("bmht,mthe->bme", att, uo) + bo                                      # (b, M, D)
            h1 = torch.relu(a @ P["score_mlp.0.weight"].T + P["score_mlp.0.bias"])
            h2 = torch.relu(h1 @ P["score_mlp.3.weight"].T + P["score_mlp.3.bias"])
            s_ = (h2 @ P["score_mlp.6.weight"].T + P["score_mlp.6.bias"]).squeeze(-1)
            out[n0:n0 + block] = torch.tanh(s_) * 0.5                                             # fusion_model.py:130
        return out

    def evaluate(self, query_embeddings, image_embeddings, target_embeddings,
                 k_values: List[int] = [1, 5, 10, 20]) -> Dict[str, float]:
        ranks = engine.matrix_rank(self.scores(query_embeddings, image_embeddings, target_embeddings))
        return _metrics_from_ranks(ranks, k_values, True, True)


class FusionModel:
    """Inference-side mirror of the reference's wrapper (`fusion_model.py:243-332`): `fusion_type` selects the head,
    `scores(query, image, target)` is its `forward` on already-normalised embeddings; `evaluate` gives the metrics
    of `evaluator_fusion.py:126-132` without the block loop."""

    TYPES = ("linear", "cross_attention", "gated", "simple_gated", "simple_gated_with_bias", "bilinear")

    def __init__(self, fusion_type: str, head):
        if fusion_type not in self.TYPES:
            raise ValueError(f"Unknown fusion type: {fusion_type}")            # fusion_model.py:287
        self.fusion_type, self.head = fusion_type, head

    def evaluate(self, query_embeddings, image_embeddings, target_embeddings, k_values: List[int] = [1, 5, 10, 20]):
        if self.fusion_type == "bilinear":
            self.head.project(image_embeddings, target_embeddings)
            return self.head.evaluate(query_embeddings, k_values)
        return self.head.evaluate(query_embeddings, image_embeddings, target_embeddings, k_values)
