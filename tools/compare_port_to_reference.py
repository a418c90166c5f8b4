"""Is the numpy restatement that `bench.py --impl reference` times a fair stand-in for the reference itself?

    python tools/compare_port_to_reference.py > profiles/r02_port_vs_reference_cpu.json    # build container only

Runs the UNMODIFIED reference function (`src/clip/eval/metrics.py:119-162`, imported from /root/reference or
$KEMR_REFERENCE -- present in the build container, absent on the GPU box, which is why bench.py itself times the port)
and the port (`bench.reference_step` -> `oracle.ref_retrieval_metrics_final(sort_kind=None)`, what `cpu_baseline` / the
reference arm execute) on the bench's own
CPU sample of C2 (250 queries x 43 000 rows x 768-d), alternating, and prints both medians and the metric dicts'
equality.  CPU only; no GPU code is touched."""
import contextlib
import io
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("KEMR_REFERENCE", "/root/reference")


def main():
    import numpy as np
    import bench
    from oracle import oracle as O
    sys.path.insert(0, REF)
    from src.clip.eval import metrics as rmetrics                       # the unmodified reference
    cfg = bench.WORKLOADS["c2"]
    s, Q, M = bench.cpu_sample(cfg, bench.REF_SAMPLE_Q)

    def ref():
        with contextlib.redirect_stdout(io.StringIO()):                 # metrics.py:147 prints the weights
            return rmetrics.compute_retrieval_metrics_final(s.query, s.target, s.image, k_values=[1, 5, 10],
                                                            t2i_weight=0.5, t2t_weight=0.5)

    def port():
        return bench.reference_step(s.query, s.image, s.target, 0.5, 0.5, cfg["k"])      # exactly what the bench times

    a, b = ref(), port()                                                # warm-up + equality
    t_ref, t_port = [], []
    for _ in range(5):
        for fn, ts in ((ref, t_ref), (port, t_port)):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
    mr, mp = statistics.median(t_ref), statistics.median(t_port)
    print(json.dumps({
        "sample": f"{Q} queries x {M} rows x {cfg['D']}-d, two galleries (the bench's cpu_baseline sample of c2)",
        "host_cores": os.cpu_count(), "numpy": np.__version__,
        "reference_unmodified": {"function": "src/clip/eval/metrics.py:119-162 compute_retrieval_metrics_final",
                                 "median_s": mr, "queries_per_s": Q / mr, "all_s": t_ref},
        "port": {"function": "bench.reference_step -> oracle.ref_retrieval_metrics_final(sort_kind=None)", "median_s": mp, "queries_per_s": Q / mp, "all_s": t_port},
        "port_over_reference_time": mp / mr,
        "metric_dicts_equal": {k: float(a[k]) == float(b[k]) for k in a},
        "same_keys": sorted(a) == sorted(b)}))


if __name__ == "__main__":
    main()
