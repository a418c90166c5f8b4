"""Where the end-to-end time of the host-buffer search goes (kemr_index_search_host through HostIndex.search):
per-call wall time of (a) the Python wrapper, (b) the bare C call with pre-built ctypes arguments, for the serving
batch (c3: 1 query, KG hits) and the bench batch (c2: 1000 queries), page-locked and pageable caller buffers.
Prints one JSON line per case.  Diagnostics, not a bench line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, fusion, index, synth  # noqa: E402


def per_call(fn, n):
    for _ in range(5):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    lib = _lib.load()
    cases = (("c3", 1, True), ("b2", 2, True), ("b4", 4, True), ("c2", 1000, False), ("q512", 512, False), ("q256", 256, False))
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    for name, Q, kg in cases:
        s = synth.make_retrieval_set(Q=Q, M=43000, D=768, seed=1, fused=True, lam=0.1, diagonal=False, with_kg=kg)
        hits_csr = None
        alpha = 1.0
        if kg:
            alpha, hits = fusion.kg_hits_for_strategy(s.kg_results, s.query_uuids, s.uuids, "weighted", {"alpha": 0.8, "sparql_weight": 0.2})
            hits_csr = (hits.rowptr.cpu().numpy(), hits.col.cpu().numpy(), hits.bonus.cpu().numpy())
        hi = index.HostIndex(synth.f32_to_bf16_bits(s.image), synth.f32_to_bf16_bits(s.target), max_queries=Q, max_k=10)
        for pinned in (True, False):
            mk = (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()) if pinned else \
                 (lambda shape, dt: torch.empty(shape, dtype=dt).numpy())
            out = (mk((Q, 10), torch.int64), mk((Q, 10), torch.float64), mk((Q,), torch.int32))
            qh = mk((Q, 768), torch.float32)
            qh[:] = s.query
            n = 200 if Q <= 4 else 50
            t_py = per_call(lambda: hi.search(qh, k=10, t2i_weight=0.5, t2t_weight=0.5, alpha=alpha, hits_csr=hits_csr, out=out), n)
            vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)   # noqa: E731
            args = (hi._h, vp(qh), Q, 0, 0.5, 0.5, float(alpha), vp(hits_csr[0]) if kg else None, vp(hits_csr[1]) if kg else None,
                    vp(hits_csr[2]) if kg else None, 10, vp(out[0]), vp(out[1]), vp(out[2]))
            t_c = per_call(lambda: lib.kemr_index_search_host(*args), n)
            rec = {"case": name, "queries": Q, "kg_hits": kg, "caller_buffers": "page-locked" if pinned else "pageable",
                   "python_wrapper_us": round(t_py, 2), "bare_c_call_us": round(t_c, 2)}
            if Q >= 256:
                qb = mk((Q, 768), torch.uint16)
                qb[:] = synth.f32_to_bf16_bits(s.query)
                rec["python_wrapper_bf16_queries_us"] = round(per_call(lambda: hi.search(qb, k=10, t2i_weight=0.5, t2t_weight=0.5, out=out), n), 2)
            print(json.dumps(rec), flush=True)
        hi.close()


if __name__ == "__main__":
    main()
