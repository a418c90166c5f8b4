"""Diagnostic: dense fp32 score matrix from the tcgen05 scan kernel vs a torch fp32 matmul, to
localise wrong tiles (rows/columns) of the CTA-pair variant.  Run on a GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knowledge_enhanced_multimodal_retrieval_b200 import engine, synth, _lib

for (Q, M, D, fused, wa, wb) in [(333, 2500, 256, True, 0.3, 0.7), (333, 2500, 256, False, 1.0, 0.0), (700, 5000, 768, True, 0.5, 0.5)]:
    s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=31, fused=fused, lam=0.15, with_kg=False, diagonal=False)
    q, img = engine.quantize(s.query), engine.quantize(s.image)
    tgt = engine.quantize(s.target) if fused else None
    ref = wa * (q.float() @ img.float().T)
    if fused:
        ref = ref + wb * (q.float() @ tgt.float().T)
    got = engine.score_matrix(q, img, tgt, wa, wb, path=_lib.PATH_MMA)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    bad = err > 1e-4
    print(f"Q={Q} M={M} D={D} fused={fused}: max err {err.max().item():.3e}, bad {int(bad.sum())} of {bad.numel()}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten().cpu().numpy()
        cols = bad.any(0).nonzero().flatten().cpu().numpy()
        print("  bad rows:", rows[:20], "... n=", len(rows), " bad cols:", cols[:20], "... n=", len(cols))
        r, c = bad.nonzero()[0].tolist()
        print("  first bad", r, c, got[r, c].item(), ref[r, c].item())
        # is the bad value a correct value of some other position?
        a = wa * (q[r].float() @ img.float().T); b = (q[r].float() @ tgt.float().T) if fused else None
        print("  img-only dot at c:", (q[r].float() @ img[c].float()).item(), " tgt-only:", (q[r].float() @ tgt[c].float()).item() if fused else None)
