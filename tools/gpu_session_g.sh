#!/bin/bash
# select-kernel A/B: warp-per-query (W=0) vs CTA-per-query with W warps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for W in 0 2 4; do
  KEMR_SELECT_WARPS=$W timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_w$W.json 2> gpurun_out/bench_c2_w$W.err; echo "W=$W rc=$?"
  python tools/benchsum.py W=$W < gpurun_out/bench_c2_w$W.json || tail -3 gpurun_out/bench_c2_w$W.err
done
