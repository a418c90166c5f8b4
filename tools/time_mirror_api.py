"""End-to-end wall time of the reference-API mirror functions on the reference's own shapes (host numpy in, python
dict out), next to the reference's numpy path (oracle.ref_*) on the host cores.  Diagnostics; prints one JSON line.
The reference side is TIMED with numpy's default argsort kind (the reference's own call, `sort_kind=None`); the metric
dicts are compared against the stable kind, the order the contract fixes.  (Records made before this distinction timed
the stable kind, which is ~3-4x slower on fp32 rows: `reference_numpy_stable_sort_ms` keeps that figure.)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O                                                      # noqa: E402
from knowledge_enhanced_multimodal_retrieval_b200 import fusion, metrics, synth     # noqa: E402


def wall(fn, reps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    res = {}
    # C1: compute_retrieval_metrics(q, g, k_values=[1,5,10]) at 4300 x 43000 x 512 (10.0 s on 8 cores)
    s = synth.make_retrieval_set(Q=4300, M=43000, D=512, seed=0, fused=False, lam=0.1, diagonal=True)
    t, got = wall(lambda: metrics.compute_retrieval_metrics(s.query, s.image, k_values=[1, 5, 10]), 5)
    t0 = time.perf_counter()
    want = O.ref_retrieval_metrics(s.query, s.image, k_values=[1, 5, 10])
    ts = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.ref_retrieval_metrics(s.query, s.image, k_values=[1, 5, 10], sort_kind=None)
    tr = time.perf_counter() - t0
    res["c1_compute_retrieval_metrics"] = {"ours_ms": t * 1e3, "reference_numpy_ms": tr * 1e3, "speedup": tr / t,
                                           "reference_numpy_stable_sort_ms": ts * 1e3,
                                           "abs_diff": {k: abs(float(got[k]) - float(want[k])) for k in want}}
    # C2: compute_retrieval_metrics_final at 1000 x 43000 x 768 and the KG-fused evaluation of evaluator.py:176-190
    s = synth.make_retrieval_set(Q=1000, M=43000, D=768, seed=1, fused=True, lam=0.1, with_kg=True, diagonal=True)
    t, got = wall(lambda: metrics.compute_retrieval_metrics_final(s.query, s.target, s.image), 5)
    t0 = time.perf_counter()
    want = O.ref_retrieval_metrics_final(s.query, s.target, s.image)
    ts = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.ref_retrieval_metrics_final(s.query, s.target, s.image, sort_kind=None)
    tr = time.perf_counter() - t0
    res["c2_compute_retrieval_metrics_final"] = {"ours_ms": t * 1e3, "reference_numpy_ms": tr * 1e3, "speedup": tr / t,
                                                 "reference_numpy_stable_sort_ms": ts * 1e3,
                                                 "abs_diff": {k: abs(float(got[k]) - float(want[k])) for k in want}}
    t, got = wall(lambda: fusion.evaluate_fused(s.query, s.target, s.image, s.kg_results, s.query_uuids, s.uuids, 0.5, 0.5,
                                                "weighted", {"alpha": 0.8, "sparql_weight": 0.2}), 3)
    t0 = time.perf_counter()
    sim = O.ref_fused_similarity(s.query, s.target, s.image, 0.5, 0.5)
    fused = O.ref_weighted_fusion(sim, s.kg_results, s.query_uuids, s.uuids, 0.8, 0.2)
    want = O.ref_metrics_from_matrix(fused)
    ts = time.perf_counter() - t0
    t0 = time.perf_counter()
    sim = O.ref_fused_similarity(s.query, s.target, s.image, 0.5, 0.5)
    O.ref_metrics_from_matrix(O.ref_weighted_fusion(sim, s.kg_results, s.query_uuids, s.uuids, 0.8, 0.2), sort_kind=None)
    tr = time.perf_counter() - t0
    res["c2_kg_fused_evaluation"] = {"ours_ms": t * 1e3, "reference_numpy_ms": tr * 1e3, "speedup": tr / t,
                                     "reference_numpy_stable_sort_ms": ts * 1e3,
                                     "abs_diff": {k: abs(float(got[k]) - float(want[k])) for k in want}}
    res["host_cores"] = os.cpu_count()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
