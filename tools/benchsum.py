"""Print a one-line summary of bench.py JSON lines read from stdin (diagnostics helper)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(" ".join(sys.argv[1:]), d["config"]["workload"][:34], "| value", round(d["value"]), "q/s | ms/step",
          round(d["ms_per_step"], 4), "| e2e", round(d["e2e"]["value"]), "|", r.get("bound"), round(r.get("achieved", 0), 1),
          r.get("unit"), "frac", round(r.get("frac", 0), 3), "kernel_ms", round(r.get("kernel_ms", 0), 4), "after", round(r.get("after_scan_ms", 0), 4), "| uncert",
          d.get("uncertified_queries"), d.get("scan_path"))
