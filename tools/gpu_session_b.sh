#!/bin/bash
# parity tests, bench lines, per-role cycle counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c2 c1 b4096}; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/bench_$w.json 2>/dev/null || tail -3 gpurun_out/bench_$w.err
  KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -1
done
