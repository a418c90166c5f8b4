#!/usr/bin/env python
"""Benchmark of the retrieval-scoring hot path (BASELINE.json metric: queries/sec, top-k=10).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2_w19|c3|c1|b64|b4096|c4|c5_b64|c5_b8192]

One "step" = one pass of the hot path over one batch of synthetic queries: similarity scan of
both galleries + weighted T2I/T2T fusion + top-k selection + canonical re-scoring.  Default
workload = BASELINE.json configs[1] ("c2": 1000 queries x 43 000 gallery x 768-d, fused
T2I+T2T, top-10).  With N > 1 (torchrun) every rank holds a replica of the 43k gallery (it is
132 MB: SURVEY section 8e reports configs 2 / 3 as replicas) and searches its own 1000-query shard
of the global batch: independent units, no data-path collective -- weak scaling over queries.
What it costs to ALSO leave every rank with every rank's results (peer-memory exchange fused into
the selection kernel, or one NCCL all-gather) is timed after the timed region and reported in
`all_results_on_every_rank`.  The path that has a real exchange step -- the row-sharded gallery of
the north_star -- is the `sharded` record of the same line.

Keys of the JSON line follow the driver's contract; see DESIGN.md §Measurement.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # The reference arm times the CPU path with all the host threads it can use.  torchrun exports OMP_NUM_THREADS=1
    # to every rank; BLAS reads these variables when numpy / torch are first imported, so they are set before that.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: Q, M, D, fused, k, kg
    "c1": dict(Q=4300, M=43000, D=512, fused=False, k=10, seed=0),
    "c2": dict(Q=1000, M=43000, D=768, fused=True, k=10, seed=1),
    "c2_w19": dict(Q=1000, M=43000, D=768, fused=True, k=10, seed=1, weights=(0.1, 0.9)),   # evaluator.py:193-194: one accumulator per gallery
    "c3": dict(Q=1, M=43000, D=768, fused=True, k=10, seed=1, kg=True),      # + KG-hit boost: alpha 0.8 / beta 0.2, ~Poisson(20) hits
    "b64": dict(Q=64, M=2_000_000, D=768, fused=False, k=10, seed=5, device_synth=True),
    "b4096": dict(Q=4096, M=1_250_000, D=768, fused=False, k=100, seed=4, device_synth=True),
    # row-sharded gallery (north_star multi-GPU path): M rows PER GPU, queries replicated, local top-k with
    # global ids, one NCCL all-gather, merge.  At 8 GPUs: c4 = 10 M rows, c5 = 100 M rows (153.6 GB).
    "c4": dict(Q=4096, M=1_250_000, D=768, fused=False, k=100, seed=4, sharded=True),
    "c5_b64": dict(Q=64, M=12_500_000, D=768, fused=False, k=10, seed=5, sharded=True),
    "c5_b8192": dict(Q=8192, M=12_500_000, D=768, fused=False, k=10, seed=5, sharded=True),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm": d["hbm_gbs"], "tensor_burst": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tensor_burst": 1590.0, "tensor_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(workload, key="scan_kernel_dram_bytes_per_launch"):
    """dram__bytes_read.sum + dram__bytes_write.sum of the scan kernel per launch, from the committed ncu --set full
    capture of this workload (profiles/ncu_traffic.json: a capture of an earlier run of the same command, not of this
    run -- a number printed under ncu is never a bench value), or None when no capture exists for it."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            return json.load(f).get(workload, {}).get(key)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference arm (CPU)
KG_ALPHA, KG_BETA = 0.8, 0.2            # RetrievalEngine defaults (retrieval.py:79)


def reference_step(q, img, tgt, wi, wt, k, kg=None):
    """The reference's own CPU path for this workload, restated in oracle/oracle.py:
    `compute_retrieval_metrics_final` (metrics.py:119-162: two fp32 sgemms, weighted sum, TWO full-row argsorts and the
    Recall@K / MRR reductions), `compute_retrieval_metrics` (metrics.py:79-116) for a single gallery, and with KG hits
    the dense indicator fusion (fusion.py:22-85) followed by `evaluate_retrieval` (fusion.py:6-20)."""
    from oracle import oracle as O
    if kg is not None:
        sim = O.ref_fused_similarity(q, tgt, img, wi, wt) if tgt is not None else O.ref_similarity(q, img)
        sim = O.ref_weighted_fusion(sim, kg[0], kg[1], kg[2], KG_ALPHA, KG_BETA)
        if sim.shape[0] > sim.shape[1]:
            raise ValueError("KG workloads are square or wide")
        return np.argsort(-sim, axis=1)[:, :k]             # serving: the ranked list (retrieval.py:74)
    # sort_kind=None: numpy's default argsort, the reference's own call (metrics.py:34,62) -- the oracle's stable kind
    # is ~4x slower on fp32 rows and would understate the reference (profiles/r02_port_vs_reference_cpu.json)
    if tgt is not None:
        return O.ref_retrieval_metrics_final(q, tgt, img, k_values=[1, 5, 10], t2i_weight=wi, t2t_weight=wt, sort_kind=None)
    return O.ref_retrieval_metrics(q, img, k_values=[1, 5, 10], sort_kind=None)


def argsort_topk_step(q, img, tgt, wi, wt, k):
    """Top-k only, the way the reference ranks (one full-row argsort, metrics.py:34)."""
    from oracle import oracle as O
    sim = O.ref_fused_similarity(q, tgt, img, wi, wt) if tgt is not None else O.ref_similarity(q, img)
    return np.argsort(-sim, axis=1)[:, :k]


def torch_topk_step(tq, timg, ttgt, wi, wt, k):
    """The fair top-k-only comparison of BASELINE.md section 5.2: torch-CPU fp32 matmuls + torch.topk."""
    import torch
    sim = wi * (tq @ timg.T) + wt * (tq @ ttgt.T) if ttgt is not None else tq @ timg.T
    return torch.topk(sim, k, dim=1)


def cpu_sample(cfg, max_q):
    from knowledge_enhanced_multimodal_retrieval_b200 import synth
    Q = min(cfg["Q"], max_q)
    M = min(cfg["M"], 43000)
    diagonal = not cfg.get("kg") and Q <= M                 # the metrics functions score query i against row i
    s = synth.make_retrieval_set(Q=Q, M=M, D=cfg["D"], seed=cfg["seed"], fused=cfg["fused"], lam=0.1,
                                 diagonal=diagonal, with_kg=bool(cfg.get("kg")))
    return s, Q, M


def kg_of(cfg, s):
    return (s.kg_results, s.query_uuids, s.uuids) if cfg.get("kg") else None


def time_cpu(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), min(ts)


REF_SAMPLE_Q = 250          # queries per CPU step: ~2-3 s of the reference's two full-row argsorts on 16 cores


def cpu_baseline_record(cfg, reps=3, max_q=REF_SAMPLE_Q):
    """Bounded sample of the workload on the host cores: the reference's function for the path, plus the two
    top-k-only variants (reference-style argsort, torch.topk).  All per-step times are scaled linearly to the full
    gallery when the sample holds fewer rows."""
    import torch
    s, Q, M = cpu_sample(cfg, max_q)
    scale = cfg["M"] / M
    wi, wt = (0.5, 0.5) if cfg["fused"] else (1.0, 0.0)
    k = cfg["k"]
    med, best = time_cpu(lambda: reference_step(s.query, s.image, s.target, wi, wt, k, kg_of(cfg, s)), reps)
    med_a, _ = time_cpu(lambda: argsort_topk_step(s.query, s.image, s.target, wi, wt, k), reps)
    tq, timg = torch.from_numpy(s.query), torch.from_numpy(s.image)
    ttgt = torch.from_numpy(s.target) if s.target is not None else None
    med_t, _ = time_cpu(lambda: torch_topk_step(tq, timg, ttgt, wi, wt, k), reps)
    what = ("dense KG-indicator fusion (fusion.py:22-85) + full-row argsort" if cfg.get("kg") else
            ("compute_retrieval_metrics_final (metrics.py:119-162): 2 sgemm + weighted sum + 2 full-row argsorts + reductions"
             if cfg["fused"] else "compute_retrieval_metrics (metrics.py:79-116): sgemm + 2 full-row argsorts + reductions"))
    return {"value": Q / (med * scale), "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{Q} queries x {M} rows x {cfg['D']}-d, {what}, numpy restatement (oracle/oracle.py), median of {reps}"
                      + ("" if scale == 1 else f", time scaled x{scale:.1f} to the full gallery"),
            "best_value": Q / (best * scale),
            "argsort_topk_only": {"value": Q / (med_a * scale), "unit": "queries/s",
                                  "what": "sgemm + weighted sum + ONE full-row argsort [:k] (round-1 definition)"},
            "torch_topk": {"value": Q / (med_t * scale), "unit": "queries/s",
                           "what": "torch-CPU fp32 matmuls + torch.topk (BASELINE.md section 5.2, the fair top-k-only figure)"},
            "blas_threads": torch.get_num_threads(), "numpy": np.__version__, "torch": torch.__version__}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    s, Q, M = cpu_sample(cfg, REF_SAMPLE_Q)
    scale = cfg["M"] / M                        # rows beyond the sample are extrapolated linearly
    wi, wt = (0.5, 0.5) if cfg["fused"] else (1.0, 0.0)
    step = lambda: reference_step(s.query, s.image, s.target, wi, wt, cfg["k"], kg_of(cfg, s))   # noqa: E731
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) * scale
    qps = Q * args.steps / dt
    base = cpu_baseline_record(cfg, reps=2)
    base["value"] = qps
    base["sample"] = f"every timed step: {base['sample'].split(', median of')[0]}"
    line = {"impl": "reference", "metric": "queries_per_sec_top%d" % cfg["k"], "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, max(1, args.gpus)),
            "cpu_baseline": base,
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, cfg, world):
    return {"workload": f"{args.workload}: {cfg['Q']} queries/GPU x {cfg['M']} gallery x {cfg['D']}-d, "
                        f"{'fused T2I+T2T (0.5/0.5)' if cfg['fused'] else 'single gallery'}, top-{cfg['k']}",
            "queries_per_step": cfg["Q"] * world, "gallery_rows": cfg["M"], "dim": cfg["D"],
            "galleries": 2 if cfg["fused"] else 1, "k": cfg["k"],
            "kg_boost": f"alpha {KG_ALPHA} / beta {KG_BETA}, ~Poisson(20) KG hits per query (CSR in the step)" if cfg.get("kg") else None,
            "parallelism": "single GPU" if world == 1 else f"query-sharded x{world}: every rank searches its own queries on its "
                                                           f"replica of the gallery (independent replicas, no collective in the step)",
            "l2": "L2 flushed (512 MiB memset) before every timed step"}


def capture_step(step, torch, dist, world):
    """Capture one step (our kernels + the NCCL all-gather) into a CUDA graph so that a multi-rank step
    is ONE launch: without it the ranks drift apart on Python launch overhead and the all-gather waits.
    Returns None (eager fallback on every rank) if any rank cannot capture."""
    g = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
    except Exception as e:                       # noqa: BLE001
        print(f"[bench] CUDA graph capture failed, running eagerly: {e}", file=sys.stderr)
        g = None
    if world > 1:
        ok = torch.tensor([1 if g is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            g = None
    return g


def timed_steps(args, torch, dist, world, lib, step, flush, use_graph):
    """W warm-up steps, a short eager pass that times the scan kernel alone (event hook in the library),
    then EXACTLY K timed steps (CUDA events, L2 flushed before each).  Returns (step_ms list, scan_ms mean)."""
    import ctypes as C
    for _ in range(max(3, args.warmup)):
        flush.zero_()
        step()
    torch.cuda.synchronize()
    n_probe = 3
    s0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_probe)]
    m0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_probe)]
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_probe)]
    for m_ in m0:
        m_.record()                      # materialise the underlying cudaEvent_t
    torch.cuda.synchronize()
    for i in range(n_probe):
        flush.zero_()
        s0[i].record()
        lib.kemr_set_scan_done_event(C.c_void_p(m0[i].cuda_event))
        step()
        e0[i].record()
    lib.kemr_set_scan_done_event(None)
    torch.cuda.synchronize()
    scan_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(s0, m0))
    timed_steps.after_scan_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(m0, e0))   # select (+ collective, merge)
    graph = capture_step(step, torch, dist, world) if use_graph else None
    run = graph.replay if graph is not None else step
    for _ in range(2):
        flush.zero_()
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    mids = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for m_ in mids:
        m_.record()
    torch.cuda.synchronize()
    for i in range(args.steps):
        flush.zero_()
        starts[i].record()
        if graph is None:                # eager launches: the scan kernel of EVERY timed step is timed in place
            lib.kemr_set_scan_done_event(C.c_void_p(mids[i].cuda_event))
        run()
        ends[i].record()
    lib.kemr_set_scan_done_event(None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if graph is None:
        scan_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(starts, mids))
        timed_steps.after_scan_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(mids, ends))
        timed_steps.scan_timing = "CUDA events around the scan kernel of every timed step"
    # phases of the fused small-batch kernel (scan + selection in one launch), stamped in the kernel with %globaltimer
    timed_steps.phases = None
    stamps = torch.zeros(3, dtype=torch.int64, device="cuda")
    lib.kemr_set_phase_stamps(C.c_void_p(stamps.data_ptr()))
    ph = []
    for _ in range(5):
        flush.zero_()
        stamps.zero_()
        step()
        torch.cuda.synchronize()
        v = stamps.cpu().tolist()
        if v[2] > v[1] > 0:
            ph.append(((v[1] - v[0]) * 1e-6, (v[2] - v[1]) * 1e-6))
    lib.kemr_set_phase_stamps(None)
    if ph:
        timed_steps.phases = {"scan_phase_ms": statistics.median(p[0] for p in ph),
                              "select_phase_ms": statistics.median(p[1] for p in ph),
                              "how": "%globaltimer stamps inside the fused kernel, 5 extra steps after the timed region"}
    else:
        timed_steps.scan_timing = "3 eager probe steps before the timed region (the timed steps are CUDA graphs)"
    return [a.elapsed_time(b) for a, b in zip(starts, ends)], scan_ms, graph is not None


# ----------------------------------------------------------------------------- our arm, row-sharded gallery
def time_plan(torch, dist, world, lib, plan, q, steps, probe=3):
    """Device-timed steps of a prepared sharded search (distributed.SearchPlan): `steps` replays bracketed by CUDA
    events (max over ranks), plus `probe` eager steps that time the scan kernel alone through the library's event hook.
    The gallery shard is far larger than L2, so every step streams it from HBM (no flush needed)."""
    import ctypes as C
    for _ in range(3):
        plan.run(q)
    torch.cuda.synchronize()
    s0 = [torch.cuda.Event(enable_timing=True) for _ in range(probe)]
    m0 = [torch.cuda.Event(enable_timing=True) for _ in range(probe)]
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(probe)]
    for m_ in m0:
        m_.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for i in range(probe):
        s0[i].record()
        lib.kemr_set_scan_done_event(C.c_void_p(m0[i].cuda_event))
        plan._step()                                     # eager launches of the same step
        e0[i].record()
    lib.kemr_set_scan_done_event(None)
    torch.cuda.synchronize()
    scan_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(s0, m0))
    after_ms = statistics.mean(a.elapsed_time(b) for a, b in zip(m0, e0))
    if world > 1:
        dist.barrier()
    st = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        st[i].record()
        plan.run(None)
        en[i].record()
    torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in zip(st, en))
    t = torch.tensor([total, scan_ms, after_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0].item()) / steps, float(t[1].item()), float(t[2].item())


def roofline_of(pk, Q, rows, D, G, scan_ms):
    scan_t = scan_ms * 1e-3
    flops, bytes_ = 2.0 * Q * G * rows * D, float(G) * rows * D * 2
    ridge = pk["tensor_burst"] * 1e12 / (pk["hbm"] * 1e9)
    long_kernel = scan_t > 2e-3          # sustained clocks apply to multi-millisecond kernels
    if Q >= ridge:
        roof = {"bound": "tensor", "achieved": flops / scan_t / 1e12,
                "peak": pk["tensor_sustained"] if long_kernel else pk["tensor_burst"], "unit": "TFLOP/s",
                "peak_kind": "sustained" if long_kernel else "burst"}
    else:
        roof = {"bound": "hbm", "achieved": bytes_ / scan_t / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["peak_source"] = pk["source"]
    roof["kernel_ms"] = scan_ms
    return roof


def sharded_records(torch, dist, world, rank, lib, pk, M_total, D, k, batches, seed=4):
    """The north_star multi-GPU layout through the product API (`distributed.ShardedGallery.plan`): a FIXED gallery of
    M_total rows cut row-wise over the ranks (strong scaling), queries replicated, local top-k with global ids, result
    exchange fused into the selection kernel over NVLink peer memory (and, for comparison, ONE NCCL all-gather), merge.
    Returns one record per query batch (rank 0; None elsewhere)."""
    from knowledge_enhanced_multimodal_retrieval_b200 import engine
    from knowledge_enhanced_multimodal_retrieval_b200.distributed import CudaLocal, ShardedGallery, shard_bounds
    lo, hi = shard_bounds(M_total, world, rank)
    gal = engine.synth_rows(hi - lo, D, seed, row_base=lo)
    sg = ShardedGallery(CudaLocal(gal), M_total)
    out = []
    for B, steps in batches:
        g = torch.Generator().manual_seed(1234 + B)                       # identical queries on every rank
        q = engine.quantize(torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=1).cuda())
        rec = {"queries_per_step": B, "gallery_rows": M_total, "rows_per_gpu": hi - lo, "dim": D, "k": k, "steps": steps,
               "scaling": "strong", "n_gpus": world}
        results = {}
        for exchange in (("peer", "nccl") if world > 1 else ("peer",)):
            plan = sg.plan(Q=B, k=k, exchange=exchange, graph=True)
            ms, scan_ms, after_ms = time_plan(torch, dist, world, lib, plan, q, steps)
            idx, score = plan.run(None)
            results[exchange] = (idx.clone(), score.clone())
            r = {"ms_per_step": ms, "value": B / (ms * 1e-3), "scan_kernel_ms": scan_ms, "after_scan_ms": after_ms,
                 "launch": "one CUDA graph per step" if plan.graph is not None else "eager launches",
                 "uncertified_queries": plan.uncertified()}
            rec[exchange] = r
            plan.close()
        if world > 1:
            rec["peer_equals_nccl"] = bool(torch.equal(results["peer"][0], results["nccl"][0]) and
                                           torch.equal(results["peer"][1], results["nccl"][1]))
        best = rec["peer"]
        rec.update({"value": best["value"], "unit": "queries/s", "ms_per_step": best["ms_per_step"],
                    "exchange": "NVLink peer memory, fused into the selection kernel (kemr_peer_*)" if world > 1 else "none (one rank)",
                    "roofline": roofline_of(pk, B, hi - lo, D, 1, best["scan_kernel_ms"]),
                    "exchange_and_merge_ms": best["after_scan_ms"]})
        rec["roofline"]["kernel"] = "scan kernel of the rank's shard (max over ranks), timed in 3 eager probe steps"
        out.append(rec)
        del q
    del sg, gal
    torch.cuda.empty_cache()
    return out if rank == 0 else None


def run_sharded(args, cfg):
    """Builder diagnostics: `--workload c4|c5_b64|c5_b8192` = cfg['M'] rows PER GPU (weak scaling in gallery size; at 8
    GPUs c4 = 10 M rows, c5 = 100 M rows), through the same product path as the `sharded` record of the bench line."""
    import torch
    import torch.distributed as dist
    from knowledge_enhanced_multimodal_retrieval_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lib = _lib.load()
    M = int(args.rows_per_gpu or cfg["M"])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    recs = sharded_records(torch, dist, world, rank, lib, peaks(), M * world, cfg["D"], cfg["k"], [(cfg["Q"], args.steps)],
                           seed=cfg["seed"])
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        r = recs[0]
        line = {"metric": "queries_per_sec_top%d" % cfg["k"], "value": r["value"], "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": 3, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {cfg['Q']} queries x {M * world} gallery rows ({M} per GPU) x {cfg['D']}-d, "
                                       f"single gallery, top-{cfg['k']}", "l2": "shard larger than L2"},
                "row_queries_per_sec": r["value"] * M * world, "roofline": r["roofline"], "sharded": r, "clocks": clocks,
                "gpu_launches": 4 * args.steps}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



def all_results_exchange(args, torch, dist, world, rank, step, flush, packed, score, idx, Q, k):
    """Not part of the bench value: the query-sharded step followed by an exchange that leaves every rank with every
    rank's top-k -- rows stored straight into every rank's buffer over NVLink peer memory by the selection kernel
    (kemr_peer_*) plus a gather kernel that waits for the flags, or ONE NCCL all-gather.  One CUDA graph per step (the
    ranks are coupled by the exchange, so launch jitter of either rank would otherwise be timed), K steps, max over
    ranks."""
    from knowledge_enhanced_multimodal_retrieval_b200.distributed import PeerExchange
    out = {"what": "query-sharded step + exchange so that every rank ends with all results (diagnostic, outside the timed region)"}
    gathered = torch.empty((world, 2, Q, k), dtype=torch.float64, device="cuda")
    g_score = torch.empty((world, Q, k), dtype=torch.float64, device="cuda")
    g_idx = torch.empty((world, Q, k), dtype=torch.int64, device="cuda")
    for name in ("peer", "nccl"):
        peer = PeerExchange(Q, k) if name == "peer" else None

        def xstep():
            if peer is not None:
                peer.begin()
            step()
            if peer is not None:
                peer.gather(Q, k, g_score, g_idx)
            else:
                dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))

        graph = capture_step(xstep, torch, dist, world)
        run = graph.replay if graph is not None else xstep
        for _ in range(3):
            flush.zero_()
            run()
        torch.cuda.synchronize()
        dist.barrier()
        st = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        en = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.zero_()
            st[i].record()
            run()
            en[i].record()
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(st, en))], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / args.steps
        ok = True
        if peer is not None:
            ok = bool(torch.equal(g_idx[rank], idx) and torch.equal(g_score[rank], score))
            peer.close()
        else:
            ok = bool(torch.equal(gathered[rank, 1].view(torch.int64), idx) and torch.equal(gathered[rank, 0], score))
        out[name] = {"ms_per_step": ms, "value": world * Q / (ms * 1e-3), "launch": "one CUDA graph per step" if graph is not None else "eager launches",
                     "own_rows_intact": ok}
    return out if rank == 0 else None


# ----------------------------------------------------------------------------- our arm (GPU)
def run_ours(args, cfg):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from knowledge_enhanced_multimodal_retrieval_b200 import _lib, engine, index, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lib = _lib.load()
    pk = peaks()
    Q, M, D, k = cfg["Q"], cfg["M"], cfg["D"], cfg["k"]
    G = 2 if cfg["fused"] else 1
    wi, wt = cfg.get("weights", (0.5, 0.5)) if cfg["fused"] else (1.0, 0.0)

    # ---- data: same gallery on every rank, per-rank query shard
    if cfg.get("device_synth"):
        img = engine.synth_rows(M, D, cfg["seed"])
        tgt = engine.synth_rows(M, D, cfg["seed"] + 1) if cfg["fused"] else None
        qsrc = img[torch.randint(0, M, (Q,), generator=torch.Generator().manual_seed(7 + rank)).cuda()]
        q_host = (qsrc.float() + 0.3 * torch.randn(Q, D, device="cuda", generator=torch.Generator("cuda").manual_seed(rank))
                  / D ** 0.5)
        q_host = torch.nn.functional.normalize(q_host, dim=1).cpu().numpy()
        host_sets = None
    else:
        s = synth.make_retrieval_set(Q=Q, M=M, D=D, seed=cfg["seed"], fused=cfg["fused"], lam=0.1, diagonal=False,
                                     with_kg=bool(cfg.get("kg")))
        if rank:
            s.query = synth.make_queries((s.image, s.target) if cfg["fused"] else (s.image,),
                                         s.target_idx, 0.1, cfg["seed"] + 100 + rank)
        img, tgt = engine.quantize(s.image), (engine.quantize(s.target) if cfg["fused"] else None)
        q_host = s.query
        host_sets = s
    q = engine.quantize(q_host)
    k_sel = engine.default_k_sel(k)
    alpha, hits, hits_csr = 1.0, None, None
    if cfg.get("kg"):
        # knowledge-graph boost in the step (BASELINE config 3): final = 0.8 * clip + 0.2 * [row in KG result]
        from knowledge_enhanced_multimodal_retrieval_b200 import fusion
        alpha, hits = fusion.kg_hits_for_strategy(host_sets.kg_results, host_sets.query_uuids, host_sets.uuids,
                                                  "weighted", {"alpha": KG_ALPHA, "sparql_weight": KG_BETA})
        hits_csr = (hits.rowptr.cpu().numpy(), hits.col.cpu().numpy(), hits.bonus.cpu().numpy())
    ws = engine.workspace_for(Q, M, D, k_sel, hits.max_per_query if hits else 0)
    flags = torch.empty((Q,), dtype=torch.int32, device="cuda")
    packed = torch.empty((2, Q, k), dtype=torch.float64, device="cuda")      # [score | idx bit-cast]: the NCCL send buffer
    score, idx = packed[0], packed[1].view(torch.int64)                        # the select kernel writes straight into it
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def step():
        engine.scan_topk_raw(q, img, tgt, wi, wt, alpha, hits, k, k_sel, engine.DEFAULT_EPS, 0, score, idx, flags, ws)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: K steps, device-timed, L2 flushed before each step
    step_ms, scan_ms_mean, graphed = timed_steps(args, torch, dist, world, lib, step, flush, False)
    scan_ms = [scan_ms_mean] * args.steps
    total_ms = sum(step_ms)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    n_uncert = int((flags & 1).sum().item())

    # ---- end to end through the C-ABI host-buffer call (gallery resident, queries/results on the host)
    hi = index.HostIndex(synth.f32_to_bf16_bits(host_sets.image) if host_sets else img.view(torch.int16).cpu().numpy().view(np.uint16),
                         None if not cfg["fused"] else (synth.f32_to_bf16_bits(host_sets.target) if host_sets
                                                        else tgt.view(torch.int16).cpu().numpy().view(np.uint16)),
                         max_queries=Q, max_k=k)
    # step inputs/outputs live in page-locked host memory (used in place by the C call)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    out = (pin((Q, k), torch.int64), pin((Q, k), torch.float64), pin((Q,), torch.int32))
    qh = pin((Q, D), torch.float32)
    qh[:] = np.ascontiguousarray(q_host, dtype=np.float32)
    for _ in range(3):
        hi.search(qh, k=k, t2i_weight=wi, t2t_weight=wt, alpha=alpha, hits_csr=hits_csr, out=out)
    if world > 1:
        dist.barrier()
    # serving-size batches take ~60 us per call: K of them are over in a millisecond and one scheduler hiccup on the host
    # would decide the figure, so they are timed over 10 K calls
    e2e_steps = args.steps * (10 if Q <= 64 else 1)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hi.search(qh, k=k, t2i_weight=wi, t2t_weight=wt, alpha=alpha, hits_csr=hits_csr, out=out)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.array_equal(out[0], idx.cpu().numpy()), "host-buffer path disagrees with the device path"
    # ---- the same calls kept two deep (kemr_index_submit_host / kemr_index_wait on two lanes of the index): step i+1's
    # queries cross PCIe while step i is scanned.  Every step still brings its own inputs from page-locked host memory
    # and returns its results to host arrays; each lane has its own buffers.
    lanes = [hi, hi.lane()]
    bufs = [(qh, out), (pin((Q, D), torch.float32), (pin((Q, k), torch.int64), pin((Q, k), torch.float64), pin((Q,), torch.int32)))]
    bufs[1][0][:] = qh

    def pipelined(n):
        for i in range(n):
            L, (qb, ob) = lanes[i & 1], bufs[i & 1]
            if i >= 2:
                L.wait()
            L.submit(qb, k=k, t2i_weight=wi, t2t_weight=wt, alpha=alpha, hits_csr=hits_csr, out=ob)
        for L in lanes[:min(n, 2)]:
            L.wait()

    pipelined(4)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pipelined(e2e_steps)
    pipe_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([pipe_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_s = float(t.item())
    assert np.array_equal(bufs[1][1][0], idx.cpu().numpy()) and np.array_equal(bufs[0][1][0], idx.cpu().numpy()), \
        "pipelined host-buffer path disagrees with the device path"
    hi.close()
    # ---- N > 1, outside the timed region: the same step when every rank must also END with every rank's results
    exchange_rec = all_results_exchange(args, torch, dist, world, rank, step, flush, packed, score, idx, Q, k) if world > 1 else None
    clocks = sampler.stop() if rank == 0 else None

    # ---- the north_star multi-GPU layout beside the headline: a fixed 10 M x 768 gallery, row-sharded over the ranks
    sharded = None
    if args.workload == "c2" and not args.no_sharded:
        del img, tgt, q, flush
        torch.cuda.empty_cache()
        recs = sharded_records(torch, dist, world, rank, lib, pk, args.sharded_rows, 768, 10, [(4096, 5), (64, 20)])
        if rank == 0:
            sharded = {"what": "fixed gallery row-sharded over the GPUs (strong scaling), queries replicated, local top-10 "
                               "with global ids, result exchange fused into the selection kernel over NVLink peer memory, "
                               "merge on every rank; through distributed.ShardedGallery.plan (one CUDA graph per step)",
                       "batches": recs}

    if rank == 0:
        info = engine.device_info()
        scan_t = statistics.mean(scan_ms) * 1e-3
        flops = 2.0 * Q * G * M * D
        bytes_ = float(G) * M * D * 2
        ridge = pk["tensor_burst"] * 1e12 / (pk["hbm"] * 1e9)
        long_kernel = scan_t > 2e-3          # sustained clocks apply to multi-millisecond kernels
        if Q >= ridge:
            roof = {"bound": "tensor", "achieved": flops / scan_t / 1e12,
                    "peak": pk["tensor_sustained"] if long_kernel else pk["tensor_burst"], "unit": "TFLOP/s",
                    "peak_kind": "sustained" if long_kernel else "burst"}
        else:
            roof = {"bound": "hbm", "achieved": bytes_ / scan_t / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["traffic"] = ncu_traffic(args.workload)
        roof["traffic_source"] = ncu_traffic(args.workload, "source")
        roof["algorithmic_bytes"] = bytes_ + Q * D * 2.0
        roof["peak_source"] = pk["source"]
        roof["kernel"] = "scan (first kernel of kemr_scan_topk)"
        roof["kernel_ms"] = scan_t * 1e3
        roof["kernel_share_of_step"] = sum(scan_ms) / sum(step_ms)
        roof["after_scan_ms"] = timed_steps.after_scan_ms
        roof["kernel_timing"] = timed_steps.scan_timing
        if timed_steps.phases:
            roof["kernel"] = "fused streaming search (scan + selection in one launch: scan_stream_kernel)"
            roof["phases"] = timed_steps.phases
            roof["scan_phase_frac"] = bytes_ / (timed_steps.phases["scan_phase_ms"] * 1e-3) / 1e9 / pk["hbm"]
        qps = world * Q * args.steps / (total_ms * 1e-3)
        line = {"metric": "queries_per_sec_top%d" % k, "value": qps, "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": workload_config(args, cfg, world),
                "e2e": {"value": world * Q * e2e_steps / pipe_s, "unit": "queries/s", "steps": e2e_steps,
                        "h2d_bytes_per_step": int(Q * D * 4), "d2h_bytes_per_step": int(Q * k * 16 + Q * 4),
                        "api": "kemr_index_submit_host / kemr_index_wait (HostIndex.submit / wait) on two lanes of one "
                               "resident index, two steps in flight: fp32 host queries in, top-k host arrays out, every "
                               "step; gallery resident in HBM; the page-locked step buffers are read / written in place "
                               "by the kernels over PCIe (no staging copy)",
                        "ms_per_step": pipe_s / e2e_steps * 1e3, "steps_in_flight": 2,
                        "blocking_call": {"value": world * Q * e2e_steps / e2e_s, "unit": "queries/s",
                                          "ms_per_step": e2e_s / e2e_steps * 1e3,
                                          "api": "kemr_index_search_host (HostIndex.search): one call at a time, each "
                                                 "returns when its results are in the host arrays (round 1's e2e figure)"}},
                "gpu_launches": ((1 if engine.scan_plan(Q, M, D, G, k_sel, wi == wt)["path"] != _lib.PATH_MMA else 2)
                                 ) * args.steps,
                "launch": "one CUDA graph per step" if graphed else "eager launches",
                "roofline": roof, "clocks": clocks,
                "uncertified_queries": n_uncert, "sm_count": info["sm_count"],
                "scan_path": "tcgen05" if engine.scan_plan(Q, M, D, G, k_sel, wi == wt)["path"] == _lib.PATH_MMA else "warp-dot"}
        # CPU baseline beside it: bounded sample of the same workload on the host cores
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_record(cfg)
        if exchange_rec is not None:
            line["all_results_on_every_rank"] = exchange_rec
        if sharded is not None:
            line["sharded"] = sharded
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the row-sharded 10 M-row record of the default workload")
    ap.add_argument("--sharded-rows", type=int, default=10_000_000, help="total gallery rows of the row-sharded record")
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="override the gallery shard size of the sharded workloads")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif cfg.get("sharded"):
        run_sharded(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
