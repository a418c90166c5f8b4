"""Multi-GPU check, launched by torchrun (one rank per GPU, NCCL):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/run_multigpu_check.py
Row-sharded gallery search + ranks must equal the single-GPU result bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knowledge_enhanced_multimodal_retrieval_b200 import engine, fusion, synth  # noqa: E402
from knowledge_enhanced_multimodal_retrieval_b200.distributed import CudaLocal, ShardedGallery, shard_bounds  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, world = dist.get_rank(), dist.get_world_size()
    s = synth.make_retrieval_set(Q=500, M=20011, D=768, seed=9, fused=True, lam=0.1, with_kg=True, diagonal=True)
    lo, hi = shard_bounds(s.M, world, rank)
    sg = ShardedGallery(CudaLocal(s.image[lo:hi], s.target[lo:hi]), s.M)
    alpha, hits = fusion.kg_hits_for_strategy(s.kg_results, s.query_uuids, s.uuids, "weighted",
                                              {"alpha": 0.8, "sparql_weight": 0.2})
    idx, score = sg.search(s.query, k=10, w_a=0.5, w_b=0.5, alpha=alpha, hits=hits)
    ranks = sg.rank_targets(s.query, torch.from_numpy(s.target_idx).cuda(), 0.5, 0.5, alpha, hits)
    # single-GPU reference on every rank
    q, img, tgt = engine.quantize(s.query), engine.quantize(s.image), engine.quantize(s.target)
    wi, ws = engine.scan_topk(q, img, tgt, 0.5, 0.5, alpha, hits, k=10)
    wr = engine.rank_targets(q, img, tgt, torch.from_numpy(s.target_idx).cuda(), 0.5, 0.5, alpha, hits)
    ok = torch.equal(idx, wi) and torch.equal(score, ws) and torch.equal(ranks, wr)
    # prepared search: result exchange fused into the selection kernel over NVLink peer memory, and the NCCL
    # all-gather fallback, eager and as a CUDA graph; several steps in a row (the buffers alternate by epoch)
    wi0, ws0 = engine.scan_topk(q, img, tgt, 0.5, 0.5, k=10)
    for exchange in ("peer", "nccl"):
        for graph in (False, True):
            plan = sg.plan(Q=q.shape[0], k=10, w_a=0.5, w_b=0.5, exchange=exchange, graph=graph)
            for step in range(4):
                qs = q if step % 2 == 0 else torch.flip(q, dims=[0])
                pi, ps = plan.run(qs)
                want_i, want_s = (wi0, ws0) if step % 2 == 0 else (torch.flip(wi0, dims=[0]), torch.flip(ws0, dims=[0]))
                good = torch.equal(pi, want_i) and torch.equal(ps, want_s) and plan.uncertified() == 0
                if not good:
                    print(f"rank {rank}: plan exchange={exchange} graph={graph} step {step} MISMATCH", flush=True)
                ok = ok and good
            plan.close()
    # unmerged gather over peer memory (query-sharded replicas): slot r of every rank = rank r's rows
    from knowledge_enhanced_multimodal_retrieval_b200.distributed import PeerExchange
    peer = PeerExchange(q.shape[0], 10)
    Qn = q.shape[0]
    qs = torch.roll(q, shifts=rank, dims=0).contiguous()                     # every rank searches different queries
    ws_ = engine.workspace_for(Qn, img.shape[0], q.shape[1], 16)
    for step in range(3):
        sc = torch.empty((Qn, 10), dtype=torch.float64, device="cuda")
        ix = torch.empty((Qn, 10), dtype=torch.int64, device="cuda")
        fl = torch.empty((Qn,), dtype=torch.int32, device="cuda")
        gs = torch.empty((world, Qn, 10), dtype=torch.float64, device="cuda")
        gi = torch.empty((world, Qn, 10), dtype=torch.int64, device="cuda")
        peer.begin()
        engine.scan_topk_raw(qs, img, tgt, 0.5, 0.5, 1.0, None, 10, 16, engine.DEFAULT_EPS, 0, sc, ix, fl, ws_)
        peer.gather(Qn, 10, gs, gi)
        ref_s = [torch.empty_like(sc) for _ in range(world)]
        ref_i = [torch.empty_like(ix) for _ in range(world)]
        dist.all_gather(ref_s, sc)
        dist.all_gather(ref_i, ix)
        good = torch.equal(gs, torch.stack(ref_s)) and torch.equal(gi, torch.stack(ref_i))
        if not good:
            print(f"rank {rank}: peer gather step {step} MISMATCH", flush=True)
        ok = ok and good
    peer.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multigpu check world={world}: {'OK' if flag.item() else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
