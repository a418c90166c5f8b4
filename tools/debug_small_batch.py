import sys, torch
sys.path.insert(0, '.')
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine
lib = _lib.load()
M, D = 43000, 768
ga, gb = engine.synth_rows(M, D, 11), engine.synth_rows(M, D, 12)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for B in (8, 64):
    q = engine.quantize(torch.nn.functional.normalize(torch.randn(B, D, generator=torch.Generator().manual_seed(B)), dim=1))
    ws = engine.workspace_for(B, M, D, 16)
    sc = torch.empty((B, 10), dtype=torch.float64, device="cuda"); ix = torch.empty((B, 10), dtype=torch.int64, device="cuda"); fl = torch.empty((B,), dtype=torch.int32, device="cuda")
    for _ in range(2):
        flush.zero_()
        engine.scan_topk_raw(q, ga, gb, 0.5, 0.5, 1.0, None, 10, 16, engine.DEFAULT_EPS, 0, sc, ix, fl, ws, _lib.PATH_MMA)
        torch.cuda.synchronize()
