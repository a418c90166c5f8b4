"""Phase timing of the selection stage inside the fused small-batch kernel (debug build: run with
KEMR_LIB=.../libkemr_debug.so).  BASELINE config 3: one query, 43 000 x 768-d x 2 galleries, ~20 KG hits, top-10, L2
flushed before the launch.  Prints the %globaltimer deltas between the phase boundaries of select_query."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine, fusion, synth  # noqa: E402

NAMES = ["A1 merge heads", "A2 gather + cut", "A3 prune", "B join KG hits", "C canonical re-score", "D order + write"]


def main():
    lib = _lib.load()
    s = synth.make_retrieval_set(Q=1, M=43000, D=768, seed=1, fused=True, lam=0.1, diagonal=False, with_kg=True)
    q, img, tgt = engine.quantize(s.query), engine.quantize(s.image), engine.quantize(s.target)
    alpha, hits = fusion.kg_hits_for_strategy(s.kg_results, s.query_uuids, s.uuids, "weighted", {"alpha": 0.8, "sparql_weight": 0.2})
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    k, ksel = 10, 16
    ws = engine.workspace_for(1, 43000, 768, ksel, hits.max_per_query)
    sc = torch.empty((1, k), dtype=torch.float64, device="cuda")
    ix = torch.empty((1, k), dtype=torch.int64, device="cuda")
    fl = torch.empty((1,), dtype=torch.int32, device="cuda")
    stamps = torch.zeros(3, dtype=torch.int64, device="cuda")
    out = np.zeros(16, dtype=np.int64)
    rows = []
    for use_hits in (True, False):
        for _ in range(6):
            flush.zero_()
            stamps.zero_()
            lib.kemr_set_phase_stamps(C.c_void_p(stamps.data_ptr()))
            engine.scan_topk_raw(q, img, tgt, 0.5, 0.5, alpha if use_hits else 1.0, hits if use_hits else None, k, ksel,
                                 engine.DEFAULT_EPS, 0, sc, ix, fl, ws)
            lib.kemr_set_phase_stamps(None)
            torch.cuda.synchronize()
            _lib.check(lib.kemr_debug_select_stamps(out.ctypes.data_as(C.c_void_p)))
            st = stamps.cpu().numpy()
            rows.append((use_hits, (st[1] - st[0]) / 1e3, (st[2] - st[1]) / 1e3, [(out[i + 1] - out[i]) / 1e3 for i in range(6)],
                         (out[0] - st[1]) / 1e3))
    for use_hits in (True, False):
        sel = [r for r in rows if r[0] == use_hits][2:]
        print(f"KG hits {'on ' if use_hits else 'off'} ({hits.max_per_query if use_hits else 0} hits): scan phase "
              f"{np.median([r[1] for r in sel]):.1f} us, selection {np.median([r[2] for r in sel]):.1f} us "
              f"(entry {np.median([r[4] for r in sel]):.1f} us after the last arrival)")
        for i, n in enumerate(NAMES):
            print(f"    {n:24s} {np.median([r[3][i] for r in sel]):6.2f} us")


if __name__ == "__main__":
    main()
