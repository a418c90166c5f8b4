"""Small invocation of every kernel family (diagnostics; compute-sanitizer is closed on this pool, so this is a plain run)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine, fusion, metrics, store, synth   # noqa: E402

s = synth.make_retrieval_set(Q=300, M=2500, D=128, seed=3, fused=True, lam=0.2, with_kg=True, diagonal=True)
q, img, tgt = engine.quantize(s.query), engine.quantize(s.image), engine.quantize(s.target)
alpha, hits = fusion.kg_hits_for_strategy(s.kg_results, s.query_uuids, s.uuids, "weighted", {"alpha": 0.8, "sparql_weight": 0.2})
for path in (_lib.PATH_WARP, _lib.PATH_MMA):
    engine.scan_topk(q, img, tgt, 0.5, 0.5, alpha, hits, k=10, path=path)          # merged accumulator + KG hits
    engine.scan_topk(q, img, tgt, 0.3, 0.7, k=20, path=path)                       # two accumulators
    engine.scan_topk(q[:7].contiguous(), img, None, k=100, path=path)             # single CTA, large k
    engine.rank_targets(q, img, tgt, torch.arange(300, device="cuda"), 0.5, 0.5, alpha, hits, path=path)
gate = np.random.default_rng(0).uniform(0.1, 0.9, 300)
engine.scan_topk(q, img, tgt, gate, 1 - gate, k=10)
engine.rank_targets(q, img, tgt, torch.arange(300, device="cuda"), gate, 1 - gate)
engine.score_matrix(q, img, tgt, 0.5, 0.5)
lists = store.HitLists(store.IdMap(s.uuids), s.kg_results, s.query_uuids)
lists.for_strategy("additive", {"delta": 0.5}, 500, 2000)
a, b = engine.scan_topk(q, img[:1200].contiguous(), None, k=10), engine.scan_topk(q, img[1200:].contiguous(), None, k=10, idx_base=1200)
engine.merge_topk(torch.stack([a[1], b[1]]), torch.stack([a[0], b[0]]), 10)
metrics.compute_recall_at_k(np.random.default_rng(1).normal(size=(50, 70)).astype(np.float32))
metrics.compute_grouped_retrieval_metrics(s.query[:100], s.image[:400], np.repeat(np.arange(100), 4))
torch.cuda.synchronize()
print("all_kernels_smoke: all launches completed")
