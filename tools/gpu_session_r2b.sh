#!/bin/bash
# Round 2, session B: parity of the new kernels (row-balanced tcgen05 scan with partial tiles, fused streaming search),
# bench c2 / c3 / c1, batch sweep, role cycles (debug build).
set -o pipefail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/b_pytest_gpu.log
for w in c2 c3 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_$w.json 2> gpurun_out/b_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/b_bench_$w.json 2>/dev/null || tail -3 gpurun_out/b_bench_$w.err
done
grep -o '"phases": {[^}]*}' gpurun_out/b_bench_c3.json
timeout 600 python tools/sweep_batch.py > gpurun_out/b_sweep.jsonl 2> gpurun_out/b_sweep.err; echo "sweep rc=$?"; cat gpurun_out/b_sweep.jsonl | cut -c1-260
for w in c2 c1; do
  KEMR_LIB=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -2 | tee gpurun_out/b_dbg_$w.log
done
