"""Batch sweep of the search step at BASELINE gallery sizes: for B in {1,2,4,8,16,32,64,128} time
kemr_scan_topk (CUDA events, L2 flushed before every step) with the streaming warp-dot kernel and with the tcgen05
kernel, at M = 43 000 x 768 x 2 galleries (serving size) and M = 2 000 000 x 768 (one gallery).  Prints one JSON line
per (shape, B, path) with the HBM roofline fraction of the scan kernel (G*M*D*2 bytes / kernel time / measured peak)."""
import ctypes as C
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine  # noqa: E402


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    lib = _lib.load()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    shapes = [("43k x 768 x 2 galleries", 43000, 768, 2), ("2M x 768 x 1 gallery", 2_000_000, 768, 1)]
    if len(sys.argv) > 1:
        shapes = [s for s in shapes if sys.argv[1] in s[0]]
    for name, M, D, G in shapes:
        ga = engine.synth_rows(M, D, 11)
        gb = engine.synth_rows(M, D, 12) if G == 2 else None
        for B in (1, 2, 4, 8, 16, 32, 64, 128):
            q = engine.quantize(torch.nn.functional.normalize(torch.randn(B, D, generator=torch.Generator().manual_seed(B)), dim=1))
            k, ksel = 10, engine.default_k_sel(10)
            ws = engine.workspace_for(B, M, D, ksel)
            sc = torch.empty((B, k), dtype=torch.float64, device="cuda")
            ix = torch.empty((B, k), dtype=torch.int64, device="cuda")
            fl = torch.empty((B,), dtype=torch.int32, device="cuda")
            ref = None
            for pname, path in (("warp-dot (streaming)", _lib.PATH_WARP), ("tcgen05", _lib.PATH_MMA)):
                if path == _lib.PATH_WARP and B > 16:
                    continue
                try:
                    def step():
                        engine.scan_topk_raw(q, ga, gb, 0.5 if G == 2 else 1.0, 0.5 if G == 2 else 0.0, 1.0, None, k, ksel,
                                             engine.DEFAULT_EPS, 0, sc, ix, fl, ws, path)
                    for _ in range(3):
                        flush.zero_(); step()
                    torch.cuda.synchronize()
                    n = 10
                    s0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
                    m0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
                    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
                    for m_ in m0:
                        m_.record()
                    torch.cuda.synchronize()
                    for i in range(n):
                        flush.zero_()
                        s0[i].record()
                        lib.kemr_set_scan_done_event(C.c_void_p(m0[i].cuda_event))
                        step()
                        e0[i].record()
                    lib.kemr_set_scan_done_event(None)
                    torch.cuda.synchronize()
                    scan = statistics.median(a.elapsed_time(b) for a, b in zip(s0, m0))
                    tot = statistics.median(a.elapsed_time(b) for a, b in zip(s0, e0))
                    # the same step without the timing event between the scan and the selection kernel (what a caller
                    # gets: the selection is then a programmatic dependent launch behind the scan)
                    for i in range(n):
                        flush.zero_()
                        s0[i].record()
                        step()
                        e0[i].record()
                    torch.cuda.synchronize()
                    tot_free = statistics.median(a.elapsed_time(b) for a, b in zip(s0, e0))
                    same = None
                    if ref is None:
                        ref = (ix.clone(), sc.clone())
                    else:
                        same = bool(torch.equal(ref[0], ix) and torch.equal(ref[1], sc))
                    gbs = G * M * D * 2 / (scan * 1e-3) / 1e9
                    print(json.dumps({"shape": name, "B": B, "path": pname, "scan_kernel_ms": round(scan, 5), "step_ms": round(tot, 5),
                                      "step_ms_without_scan_event": round(tot_free, 5),
                                      "hbm_GBs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peaks["hbm_gbs"], 3),
                                      "uncertified": int((fl & 1).sum().item()), "same_result_as_other_path": same}), flush=True)
                except Exception as e:      # noqa: BLE001
                    print(json.dumps({"shape": name, "B": B, "path": pname, "error": str(e)[:200]}), flush=True)
        del ga, gb
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
