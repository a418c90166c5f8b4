"""bench.py contract on a CPU-only box: the reference arm runs without a GPU and prints one JSON line with the
keys the driver reads; under a multi-rank launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_json_line():
    lines = [l for l in _run().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries_per_sec_top10" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("c2:") and d["config"]["gallery_rows"] == 43000 and d["config"]["dim"] == 768
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}).strip() == ""


def test_reference_step_sorts_like_the_reference_not_like_the_oracle(monkeypatch):
    """The timing legs must call numpy's argsort the way the reference does (`np.argsort(-S, axis=1)`, default kind:
    metrics.py:34,62) -- the oracle's stable kind is ~4x slower on fp32 rows and would understate the reference
    (profiles/r02_port_vs_reference_cpu.json).  The parity legs keep the stable kind."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import oracle as O
    kinds = []
    real = np.argsort
    monkeypatch.setattr(np, "argsort", lambda a, *args, **kw: (kinds.append(kw.get("kind")), real(a, *args, **kw))[1])
    rng = np.random.default_rng(0)
    q, img, tgt = (rng.standard_normal((n, 32), dtype=np.float32) for n in (6, 40, 40))
    timed = bench.reference_step(q, img, tgt, 0.5, 0.5, 10)
    assert kinds == [None, None]                         # two full-row argsorts, both the reference's own call
    kinds.clear()
    assert bench.reference_step(q, img, None, 1.0, 0.0, 10) == O.ref_retrieval_metrics(q, img, k_values=[1, 5, 10])
    assert kinds == [None, None, "stable", "stable"]
    assert timed == O.ref_retrieval_metrics_final(q, tgt, img, k_values=[1, 5, 10])     # tie-free data: same metrics
