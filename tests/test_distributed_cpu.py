"""CPU, world_size 2, gloo: the sharding / all-gather / merge / all-reduce logic of
distributed.ShardedGallery, with the oracle standing in for the per-rank CUDA compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from knowledge_enhanced_multimodal_retrieval_b200 import synth
from knowledge_enhanced_multimodal_retrieval_b200.distributed import ShardedGallery, shard_bounds


class OracleLocal:
    """Same interface as distributed.CudaLocal, computed by oracle/oracle.py on CPU tensors."""

    def __init__(self, image, target):
        self.image, self.target = image, target

    def prepare_queries(self, q):
        return torch.as_tensor(q)

    def _scores(self, q, w_a, w_b, alpha, hits):
        qn = q.numpy()
        sa = O.canon_dot64(qn, self.image)
        sb = O.canon_dot64(qn, self.target) if self.target is not None else None
        bonus = None
        if hits is not None:
            bonus = np.zeros_like(sa)
            rp, col, bon = hits
            for i in range(len(rp) - 1):
                bonus[i, col[rp[i]:rp[i + 1]]] = bon[rp[i]:rp[i + 1]]
        return O.canon_fused64(sa, sb, w_a, w_b, alpha, bonus)

    def topk(self, q, k, w_a, w_b, alpha, hits, idx_base):
        s = self._scores(q, w_a, w_b, alpha, hits)
        idx, sc = O.canon_topk(s, k)
        pad = k - idx.shape[1]
        if pad > 0:
            idx = np.concatenate([idx, np.full((idx.shape[0], pad), -1 - idx_base)], axis=1)
            sc = np.concatenate([sc, np.full((sc.shape[0], pad), -np.inf)], axis=1)
        gidx = np.where(idx >= 0, idx + idx_base, -1)
        return torch.from_numpy(gidx), torch.from_numpy(sc)

    def pair_scores(self, q, rows_local, w_a, w_b, alpha, bonus):
        s = self._scores(q, w_a, w_b, alpha, None)
        return torch.from_numpy(s[np.arange(s.shape[0]), rows_local.numpy()].copy())

    def count_ahead(self, q, t_score, t_gidx, w_a, w_b, alpha, hits, idx_base):
        s = self._scores(q, w_a, w_b, alpha, hits)
        t = t_score.numpy()[:, None]
        gid = idx_base + np.arange(s.shape[1])[None, :]
        ahead = (s > t) | ((s == t) & (gid < t_gidx.numpy()[:, None]))
        return torch.from_numpy(ahead.sum(axis=1).astype(np.int64))

    def merge(self, scores, idx, k):
        R, Q, _ = scores.shape
        oi = np.full((Q, k), -1, np.int64)
        os_ = np.full((Q, k), -np.inf)
        for qi in range(Q):
            cand = [(-float(scores[r, qi, j]), int(idx[r, qi, j])) for r in range(R) for j in range(k)
                    if int(idx[r, qi, j]) >= 0]
            cand.sort()
            for n, (ns, ix) in enumerate(cand[:k]):
                oi[qi, n], os_[qi, n] = ix, -ns
        return torch.from_numpy(oi), torch.from_numpy(os_)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = synth.make_retrieval_set(Q=24, M=301, D=64, seed=77, fused=True, lam=0.25, diagonal=False)
        s.image[150:160] = s.image[10:20]           # exact cross-shard ties: lowest global index must win
        s.target[150:160] = s.target[10:20]
        lo, hi = shard_bounds(s.M, world, rank)
        sg = ShardedGallery(OracleLocal(s.image[lo:hi], s.target[lo:hi]), s.M)
        idx, score = sg.search(torch.from_numpy(s.query), k=12, w_a=0.3, w_b=0.7)
        ranks = sg.rank_targets(torch.from_numpy(s.query), torch.from_numpy(s.target_idx), 0.3, 0.7)
        full = O.canon_fused64(O.canon_dot64(s.query, s.image), O.canon_dot64(s.query, s.target), 0.3, 0.7)
        widx, wscore = O.canon_topk(full, 12)
        ok = (np.array_equal(idx.numpy(), widx) and np.array_equal(score.numpy(), wscore)
              and np.array_equal(ranks.numpy(), O.canon_rank(full, s.target_idx)))
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_gallery_world2_gloo():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    for M in (1, 7, 43000, 10_000_001):
        for w in (1, 2, 3, 8):
            edges = [shard_bounds(M, w, r) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == M
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
