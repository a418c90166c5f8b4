#!/bin/bash
# Last validation of the round (one GPU): parity suite, smoke, the row-sharded metrics path at one rank, default bench line.
set -o pipefail
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z_pytest_gpu.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/z_smoke.log
timeout 90 python tools/time_sharded_metrics.py > gpurun_out/z_sharded_metrics_n1.jsonl 2> gpurun_out/z_sharded_metrics_n1.err; echo "sharded metrics rc=$?"
cat gpurun_out/z_sharded_metrics_n1.jsonl; tail -3 gpurun_out/z_sharded_metrics_n1.err
timeout 200 python bench.py > gpurun_out/z_bench_default.json 2> gpurun_out/z_bench_default.err; echo "bench default rc=$?"
python tools/benchsum.py default < gpurun_out/z_bench_default.json || tail -3 gpurun_out/z_bench_default.err
