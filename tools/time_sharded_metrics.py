"""Throughput of the row-sharded METRICS path (`distributed.ShardedGallery.rank_targets`, SURVEY 8e: all-reduce MAX of
the owner's target score + all-reduce SUM of the per-shard counts) on a fixed gallery cut row-wise over the ranks.

    python tools/time_sharded_metrics.py [--rows 10000000] [--dim 768] [--batches 4096,64] [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/time_sharded_metrics.py

The targets are each query's k-th search result, so every rank must come out as exactly k (checked on every rank).
The call is eager and synchronises with the host (certificate / overflow flags), so a step is timed by the wall clock
between a barrier + device synchronise on both sides, max over ranks.  Rank 0 prints one JSON line per batch size."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batches", default="4096,64")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from knowledge_enhanced_multimodal_retrieval_b200 import engine
    from knowledge_enhanced_multimodal_retrieval_b200.distributed import CudaLocal, ShardedGallery, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lo, hi = shard_bounds(args.rows, world, rank)
    sg = ShardedGallery(CudaLocal(engine.synth_rows(hi - lo, args.dim, 4, row_base=lo)), args.rows)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for B in (int(b) for b in args.batches.split(",")):
        g = torch.Generator().manual_seed(1234 + B)                       # identical queries on every rank
        q = engine.quantize(torch.nn.functional.normalize(torch.randn(B, args.dim, generator=g), dim=1).cuda())
        idx, _ = sg.search(q, k=args.k)
        targets = idx[:, args.k - 1].contiguous()
        ms = []
        for it in range(args.warmup + args.steps):
            fence()
            t0 = time.perf_counter()
            ranks = sg.rank_targets(q, targets)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if it >= args.warmup:
                ms.append(float(dt.item()) * 1e3)
        ok = bool((ranks == args.k).all().item())
        if rank == 0:
            step = sum(ms) / len(ms)
            print(json.dumps({"what": "ShardedGallery.rank_targets (pair scores, all-reduce MAX, count-ahead scan, "
                                      "all-reduce SUM; eager, host-synchronising), wall clock, max over ranks",
                              "n_gpus": world, "gallery_rows": args.rows, "rows_per_gpu": hi - lo, "dim": args.dim,
                              "queries_per_step": B, "steps": args.steps, "ms_per_step": step, "ms_best": min(ms),
                              "value": B / (step * 1e-3), "unit": "queries/s",
                              "every_rank_equals_k": ok}), flush=True)
        assert ok, "rank of the k-th search result is not k"
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
