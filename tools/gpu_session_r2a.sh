#!/bin/bash
# Round 2, session A: the new parity tests (full-size audits, heads / grouped ground truth, C4 / C5 shapes), baseline
# numbers of this box for c2 / c3 (with KG hits) / c1, role cycles + launch skew of the tcgen05 scan.
set -o pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/a_pytest_gpu.log
grep "audit" gpurun_out/a_pytest_gpu.log
for w in c2 c3 c1; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_$w.json 2> gpurun_out/a_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/a_bench_$w.json 2>/dev/null || tail -3 gpurun_out/a_bench_$w.err
done
for w in c2 c1; do
  KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -2 | tee gpurun_out/a_dbg_$w.log
done
