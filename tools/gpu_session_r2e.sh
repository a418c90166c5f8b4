#!/bin/bash
# Round 2, session E: streaming floor microbenchmark; parity (new heads, InfoNCE); bench c2 / c1 with the per-mode
# kernels; role cycles.
set -o pipefail
mkdir -p gpurun_out
./tools/microbench/stream_floor | tee gpurun_out/e_stream_floor.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/e_pytest_gpu.log
for w in c2 c1; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/e_bench_$w.json 2> gpurun_out/e_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/e_bench_$w.json 2>/dev/null || tail -3 gpurun_out/e_bench_$w.err
done
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
for w in c2 c1; do
  KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-sharded > /dev/null 2> gpurun_out/e_dbg_$w.err
  grep "kemr mma dbg" gpurun_out/e_dbg_$w.err | head -2 | cut -c1-420 || tail -3 gpurun_out/e_dbg_$w.err
done
