"""Parity of the tcgen05 scan under a forced kernel variant (cluster size, stage layout, accumulator layout).
The variant is chosen by KEMR_MMA_* environment variables that libkemr reads once, so every variant needs its own
process: tests/test_gpu_variants.py launches this script.  Exit code 0 = every case bit-exact vs the oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import coracle as CO                                                 # noqa: E402
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine, synth     # noqa: E402

# Q, M, D, galleries, weights, k
CASES = [
    (600, 5000, 768, 2, (0.5, 0.5), 10),      # equal weights: merged accumulator (double stage with CTA pairs)
    (600, 5000, 768, 2, (0.1, 0.9), 10),      # two accumulators
    (1100, 3000, 512, 1, (1.0, 0.0), 10),     # single gallery, odd number of 256-query blocks (phantom block in a quad)
    (520, 7000, 256, 1, (1.0, 0.0), 100),     # large k: many short lists
    (300, 2500, 128, 2, (0.5, 0.5), 20),
    (3300, 3400, 64, 1, (1.0, 0.0), 10),      # more query blocks than units can split evenly: flattened (block, tile) ranges
    (3300, 3400, 64, 2, (0.3, 0.7), 10),
]


def main():
    bad = 0
    for Q, M, D, G, w, k in CASES:
        s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=7 * Q + M, fused=G == 2, lam=0.2, diagonal=True)
        q, a = engine.quantize(s.query), engine.quantize(s.image)
        b = engine.quantize(s.target) if G == 2 else None
        k_sel = engine.default_k_sel(k)
        if engine.scan_plan(Q, M, D, G, k_sel, w[0] == w[1])["path"] != _lib.PATH_MMA:
            print(f"case {(Q, M, D, G, w, k)}: planner did not choose the tcgen05 kernel", file=sys.stderr)
            bad += 1
            continue
        idx, score = engine.scan_topk(q, a, b, w[0], w[1], k=k, path=_lib.PATH_MMA)
        can = CO.scores(s.query, s.image, s.target if G == 2 else None, w[0], w[1])       # canonical binary64, C oracle
        widx, wscore, wrank = CO.topk_rank(s.query, s.image, s.target if G == 2 else None, w[0], w[1], k=k,
                                           target=s.target_idx)
        ranks = engine.rank_targets(q, a, b, torch.from_numpy(s.target_idx).cuda(), w[0], w[1], path=_lib.PATH_MMA)
        dense = engine.score_matrix(q, a, b, w[0], w[1], path=_lib.PATH_MMA).cpu().numpy()
        ok = (np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(score.cpu().numpy(), wscore)
              and np.array_equal(ranks.cpu().numpy(), wrank) and int((engine.last_flags() != 0).sum()) == 0
              and float(np.abs(dense - can).max()) < engine.DEFAULT_EPS / 4)
        print(f"case Q={Q} M={M} D={D} G={G} w={w} k={k}: {'ok' if ok else 'MISMATCH'}")
        bad += 0 if ok else 1
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
