"""Row-sharded gallery across the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink).  The gallery is cut into contiguous
row ranges, one per rank; queries are replicated.  A search is: local fused scan + top-k with
GLOBAL row ids -> ONE all-gather of (score f64, idx i64) x k x Q per rank -> merge by (score desc,
index asc) on every rank.  Ranks for Recall@K/MRR: the rank owning a query's target row computes
its canonical score (all-reduce MAX, others contribute -inf), every rank counts the rows of its
shard that outrank it, and one all-reduce(SUM) of int64[Q] gives rank-1.

The reference has no counterpart (its scoring is single-process numpy); this module only
distributes calls whose single-GPU form is already parity-checked.  The local compute is
injectable so that the collective logic can be exercised with the gloo backend on CPUs
(tests/test_distributed_cpu.py passes the oracle); the default is the CUDA engine.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(M: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous rows [lo, hi) of `rank`; rank order == index order, so lowest-index ties survive the merge."""
    return M * rank // world, M * (rank + 1) // world


class CudaLocal:
    """Local compute on this rank's shard through libkemr.so."""

    def __init__(self, image_shard, target_shard=None):
        from . import engine
        self.e = engine
        self.image = engine.quantize(image_shard)
        self.target = engine.quantize(target_shard) if target_shard is not None else None
        self.device = self.image.device

    def prepare_queries(self, q):
        return self.e.quantize(q)

    def topk(self, q, k, w_a, w_b, alpha, hits, idx_base):
        return self.e.scan_topk(q, self.image, self.target, w_a, w_b, alpha, hits, k, idx_base=idx_base)

    def pair_scores(self, q, rows_local, w_a, w_b, alpha, bonus):
        n = q.shape[0]
        return self.e.score_pairs(q, self.image, self.target, torch.arange(n, device=q.device), rows_local,
                                  w_a, w_b, alpha, bonus)

    def count_ahead(self, q, t_score, t_gidx, w_a, w_b, alpha, hits, idx_base):
        return self.e.rank_count(q, self.image, self.target, t_score, t_gidx, w_a, w_b, alpha, hits,
                                 idx_base=idx_base)

    def merge(self, scores, idx, k):
        return self.e.merge_topk(scores, idx, k)


class ShardedGallery:
    def __init__(self, local, M_total: int, group=None):
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.M_total = M_total
        self.lo, self.hi = shard_bounds(M_total, self.world, self.rank)

    # ---- top-k
    def search(self, q, k: int = 10, w_a: float = 1.0, w_b: float = 0.0, alpha: float = 1.0, hits=None):
        """hits: KGHits over GLOBAL rows (or None); returns global (idx [Q,k], score [Q,k]) on every rank."""
        q = self.local.prepare_queries(q)
        local_hits = hits.shard(self.lo, self.hi) if hits is not None else None
        idx, score = self.local.topk(q, k, w_a, w_b, alpha, local_hits, self.lo)
        if self.world == 1:
            return idx, score
        Q = idx.shape[0]
        packed = torch.empty((Q, 2 * k), dtype=torch.float64, device=idx.device)
        packed[:, :k] = score
        packed[:, k:] = idx.contiguous().view(torch.float64)          # bit-cast, exact
        gathered = torch.empty((self.world, Q, 2 * k), dtype=torch.float64, device=idx.device)
        dist.all_gather_into_tensor(gathered.view(self.world * Q, 2 * k), packed, group=self.group)
        g_score = gathered[:, :, :k].contiguous()
        g_idx = gathered[:, :, k:].contiguous().view(torch.int64)
        return self.local.merge(g_score, g_idx, k)

    # ---- ranks of target rows (global ids)
    def rank_targets(self, q, target_gidx: torch.Tensor, w_a: float = 1.0, w_b: float = 0.0, alpha: float = 1.0,
                     hits=None):
        q = self.local.prepare_queries(q)
        dev = q.device if isinstance(q, torch.Tensor) else target_gidx.device
        tg = target_gidx.to(device=dev, dtype=torch.int64)
        mine = (tg >= self.lo) & (tg < self.hi)
        local_hits = hits.shard(self.lo, self.hi) if hits is not None else None
        bonus = None
        if hits is not None:
            from .engine import target_bonus
            bonus = target_bonus(local_hits, torch.where(mine, tg - self.lo, torch.full_like(tg, -1)))
        rows = torch.where(mine, tg - self.lo, torch.zeros_like(tg))
        t = self.local.pair_scores(q, rows, w_a, w_b, alpha, bonus)
        t = torch.where(mine, t, torch.full_like(t, float("-inf")))
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)    # exactly one owner per target
        count = self.local.count_ahead(q, t, tg, w_a, w_b, alpha, local_hits, self.lo)
        if self.world > 1:
            dist.all_reduce(count, op=dist.ReduceOp.SUM, group=self.group)
        return count + 1
