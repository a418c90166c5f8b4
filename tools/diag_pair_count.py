"""Diagnostic: rank counts from the tcgen05 scan (count mode) vs the oracle; prints which queries differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from knowledge_enhanced_multimodal_retrieval_b200 import engine, synth, _lib

cases = [(333, 2500, 256, True), (130, 128, 64, False), (130, 300, 64, False), (64, 128, 64, False), (130, 128, 128, False)]
for (Q, M, D, fused) in cases:
    s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=31, fused=fused, lam=0.15, with_kg=False, diagonal=False)
    q, img = engine.quantize(s.query), engine.quantize(s.image)
    tgt = engine.quantize(s.target) if fused else None
    si = O.canon_dot64(s.query, s.image)
    st = O.canon_dot64(s.query, s.target) if fused else None
    wa, wb = (0.5, 0.5) if fused else (1.0, 0.0)
    can = O.canon_fused64(si, st, wa, wb)
    want = O.canon_rank(can, s.target_idx)
    tidx = torch.from_numpy(s.target_idx).cuda()
    got = engine.rank_targets(q, img, tgt, tidx, wa, wb, path=_lib.PATH_MMA).cpu().numpy()
    got2 = engine.rank_targets(q, img, tgt, tidx, wa, wb, path=_lib.PATH_MMA).cpu().numpy()
    d = got - want
    bad = np.nonzero(d)[0]
    print(f"PAIR={os.environ.get('KEMR_MMA_PAIR')} Q={Q} M={M} D={D} fused={fused}: {len(bad)} of {Q} ranks differ; hist {dict(zip(*np.unique(d, return_counts=True)))}; repeatable={np.array_equal(got, got2)}")
    if len(bad):
        print("  bad queries:", bad[:24], "...", bad[-6:])
        dense = engine.score_matrix(q, img, tgt, wa, wb, path=_lib.PATH_MMA).cpu().numpy().astype(np.float64)
        t = can[np.arange(Q), s.target_idx]
        eps = 2e-5 * (1 + 1 / 64)
        for i in bad[:6]:
            above = int((dense[i] > t[i] + eps).sum()); inband = int((np.abs(dense[i] - t[i]) <= eps).sum())
            print(f"   q{i}: got {got[i]} want {want[i]} target {s.target_idx[i]} t={t[i]:.8f} dense@t={dense[i, s.target_idx[i]]:.8f} "
                  f"dense-above-band={above} in-band={inband} canon-ahead={int((can[i] > t[i]).sum())}")
