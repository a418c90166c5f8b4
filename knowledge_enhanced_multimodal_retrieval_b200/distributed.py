"""Row-sharded gallery across the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink).  The gallery is cut into contiguous
row ranges, one per rank; queries are replicated.  A search is: local fused scan + top-k with
GLOBAL row ids -> exchange of (score f64, idx i64) x k x Q per rank -> merge by (score desc,
index asc) on every rank.  Two exchanges:

* `"peer"` (default on a CUDA box): FUSED into the selection kernel -- every rank's selection kernel stores its k
  result rows straight into slot r of every rank's exchange buffer over NVLink peer memory and releases a per-query
  flag there; the merge kernel waits for the flags in local memory (`kemr_peer_*`).  No collective launch, no host
  synchronisation; the whole step is three kernels and can be replayed as one CUDA graph.
* `"nccl"`: ONE `all_gather_into_tensor` of the packed rows, then `kemr_merge_topk` (the checked fallback; the two
  are bit-identical, tests/run_multigpu_check.py).

Ranks for Recall@K/MRR: the rank owning a query's target row computes its canonical score (all-reduce MAX, others
contribute -inf), every rank counts the rows of its shard that outrank it, and one all-reduce(SUM) of int64[Q] gives
rank-1.

The reference has no counterpart (its scoring is single-process numpy); this module only
distributes calls whose single-GPU form is already parity-checked.  The local compute is
injectable so that the collective logic can be exercised with the gloo backend on CPUs
(tests/test_distributed_cpu.py passes the oracle); the default is the CUDA engine.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(M: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous rows [lo, hi) of `rank`; rank order == index order, so lowest-index ties survive the merge."""
    return M * rank // world, M * (rank + 1) // world


class PeerExchange:
    """Exchange buffers of `kemr_peer_*`, one per rank, mapped into every rank (CUDA IPC).  Collective constructor:
    every rank of `group` must create it with the same limits."""

    def __init__(self, max_queries: int, max_k: int, group=None):
        from . import _lib
        self._lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.max_queries, self.max_k = int(max_queries), int(max_k)
        h = C.c_void_p()
        mine = (C.c_ubyte * 64)()
        _lib.check(self._lib.kemr_peer_create(self.rank, self.world, self.max_queries, self.max_k, C.byref(h), mine))
        self._h = h
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine), group=group)
            blob = b"".join(handles)
            _lib.check(self._lib.kemr_peer_connect(self._h, C.c_char_p(blob)))
            dist.barrier(group=group)                                  # every rank has mapped every buffer

    def begin(self):
        from . import _lib
        _lib.check(self._lib.kemr_peer_begin(self._h, torch.cuda.current_stream().cuda_stream))

    def merge(self, Q: int, k: int, out_score: torch.Tensor, out_idx: torch.Tensor):
        from . import _lib
        _lib.check(self._lib.kemr_peer_merge(self._h, int(Q), int(k), C.c_void_p(out_score.data_ptr()),
                                             C.c_void_p(out_idx.data_ptr()), torch.cuda.current_stream().cuda_stream))

    def gather(self, Q: int, k: int, out_score: torch.Tensor, out_idx: torch.Tensor):
        """out_*[rank][Q][k]: every rank's own rows, unmerged (query-sharded replicas)."""
        from . import _lib
        _lib.check(self._lib.kemr_peer_gather(self._h, int(Q), int(k), C.c_void_p(out_score.data_ptr()),
                                              C.c_void_p(out_idx.data_ptr()), torch.cuda.current_stream().cuda_stream))

    def close(self):
        if getattr(self, "_h", None):
            if self.world > 1 and dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier(group=self.group)                         # nobody still stores into a buffer about to go
            self._lib.kemr_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.world == 1:
                self.close()
        except Exception:
            pass


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
        """hits: KGHits over GLOBAL rows (or None); returns global (idx [Q,k], score [Q,k]) on every rank.
        Certified search (uncertified queries are re-run with a wider margin) + ONE all-gather + merge."""
        q = self.local.prepare_queries(q)
        local_hits = hits.shard(self.lo, self.hi) if hits is not None else None
        idx, score = self.local.topk(q, k, w_a, w_b, alpha, local_hits, self.lo)
        if self.world == 1:
            return idx, score
        Q = idx.shape[0]
        packed = torch.empty((2, Q, k), dtype=torch.float64, device=idx.device)
        packed[0] = score
        packed[1] = idx.contiguous().view(torch.float64)              # bit-cast, exact
        gathered = torch.empty((self.world, 2, Q, k), dtype=torch.float64, device=idx.device)
        dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1), group=self.group)
        return self.local.merge(gathered[:, 0], gathered[:, 1].contiguous().view(torch.int64), k)

    def plan(self, Q: int, k: int = 10, w_a: float = 1.0, w_b: float = 0.0, alpha: float = 1.0,
             exchange: str = "peer", graph: bool = True) -> "SearchPlan":
        """A prepared search of fixed shape for the serving loop: see `SearchPlan` (CUDA engine only)."""
        return SearchPlan(self, Q, k, w_a, w_b, alpha, exchange, graph)

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


class SearchPlan:
    """Prepared row-sharded search of a fixed shape (Q queries, top-k, fusion weights): buffers allocated once, no host
    synchronisation in the step, the step captured as ONE CUDA graph.  Collective: every rank builds the same plan
    and calls `run` the same number of times.

        plan = ShardedGallery(CudaLocal(shard), M_total).plan(Q=4096, k=10)
        idx, score = plan.run(queries_bf16)            # device tensors, global row ids, valid until the next run

    exchange = "peer": the selection kernel stores its rows into every rank's exchange buffer over NVLink and the
    merge kernel waits on per-query flags (kemr_peer_*); "nccl": all_gather_into_tensor + kemr_merge_topk.
    Certificates are not re-run here: `uncertified()` reports how many local queries would need `ShardedGallery.search`.
    """

    def __init__(self, sg: ShardedGallery, Q: int, k: int, w_a, w_b, alpha, exchange: str, graph: bool):
        from . import engine
        if not isinstance(sg.local, CudaLocal):
            raise engine.KemrError("SearchPlan needs the CUDA engine (CudaLocal)")
        if exchange not in ("peer", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.sg, self.e = sg, engine
        self.Q, self.k, self.w_a, self.w_b, self.alpha = int(Q), int(k), float(w_a), float(w_b), float(alpha)
        self.exchange = exchange if sg.world > 1 or exchange == "peer" else "nccl"
        loc = sg.local
        dev = loc.device
        D = loc.image.shape[1]
        self.k_sel = engine.default_k_sel(self.k)
        # selection margin: queries are L2-normalised by contract (the plan never reads them back), the shard's row
        # norms and the weights scale it (engine.eps_for)
        g_norm = engine.row_norm_max(loc.image) * abs(self.w_a)
        if loc.target is not None:
            g_norm += engine.row_norm_max(loc.target) * abs(self.w_b)
        self.eps = engine.DEFAULT_EPS * max(1.0, g_norm * (1.0 + 2.0 ** -7))
        self.ws = torch.empty(int(engine._lib.load().kemr_workspace_bytes(self.Q, loc.image.shape[0], D, self.k_sel, 0)),
                              dtype=torch.uint8, device=dev)                # private workspace: plans may overlap other calls
        self.q = torch.empty((self.Q, D), dtype=torch.bfloat16, device=dev)
        self.flags = torch.empty((self.Q,), dtype=torch.int32, device=dev)
        self.packed = torch.empty((2, self.Q, self.k), dtype=torch.float64, device=dev)      # [score | idx bit-cast]
        self.score, self.idx = self.packed[0], self.packed[1].view(torch.int64)
        self.out_score = torch.empty((self.Q, self.k), dtype=torch.float64, device=dev)
        self.out_idx = torch.empty((self.Q, self.k), dtype=torch.int64, device=dev)
        self.gathered = None
        self.peer = None
        if self.exchange == "peer":
            self.peer = PeerExchange(self.Q, self.k, sg.group)
        elif sg.world > 1:
            self.gathered = torch.empty((sg.world, 2, self.Q, self.k), dtype=torch.float64, device=dev)
        self.graph = None
        if graph:
            self._capture()

    # one step, enqueued on the current stream
    def _step(self):
        sg, loc, e = self.sg, self.sg.local, self.e
        if self.peer is not None:
            self.peer.begin()
        e.scan_topk_raw(self.q, loc.image, loc.target, self.w_a, self.w_b, self.alpha, None, self.k, self.k_sel,
                        self.eps, sg.lo, self.score, self.idx, self.flags, self.ws)
        if self.peer is not None:
            self.peer.merge(self.Q, self.k, self.out_score, self.out_idx)
        elif sg.world > 1:
            dist.all_gather_into_tensor(self.gathered.view(-1), self.packed.view(-1), group=sg.group)
            # gathered is [rank][score | idx][Q][k]: both bases point into it, one rank is 2*Q*k entries further on
            e._lib.check(e._lib.load().kemr_merge_topk_strided(
                C.c_void_p(self.gathered.data_ptr()), C.c_void_p(self.gathered[0, 1].data_ptr()), 2 * self.Q * self.k,
                sg.world, self.Q, self.k, C.c_void_p(self.out_score.data_ptr()), C.c_void_p(self.out_idx.data_ptr()),
                torch.cuda.current_stream().cuda_stream))
        else:
            self.out_score.copy_(self.score)
            self.out_idx.copy_(self.idx)

    def _capture(self):
        """Two eager steps on a side stream (lazy initialisation, NCCL warm-up), then capture.  Falls back to eager
        launches on every rank if any rank cannot capture."""
        g = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._step()
                self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()
        except Exception:                                # noqa: BLE001
            g = None
            torch.cuda.synchronize()
        if self.sg.world > 1:
            ok = torch.tensor([1 if g is not None else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.sg.group)
            if int(ok.item()) == 0:
                g = None
        self.graph = g

    def run(self, q_bf16: Optional[torch.Tensor] = None):
        """One search step.  `q_bf16` (bf16 [Q, D] CUDA tensor) is copied into the plan's query buffer; None re-runs
        the queries already there.  Returns (idx int64 [Q,k], score f64 [Q,k]): global ids, identical on every rank."""
        if q_bf16 is not None:
            self.q.copy_(q_bf16, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.out_idx, self.out_score

    def uncertified(self) -> int:
        return int((self.flags & 1).sum().item())

    def close(self):
        self.graph = None
        if self.peer is not None:
            self.peer.close()
            self.peer = None
