#!/bin/bash
# Round 2, session F: parity; streaming kernel v3 (LDG, 2 CTAs / SM) on c3 + batch sweep; tcgen05 with half-weight
# remainder units and the L2 prefetch of the next tile (on / off) on c2 / c1; role cycles.
set -o pipefail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/f_pytest_gpu.log
for w in c3 c2 c1; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/f_bench_$w.json 2> gpurun_out/f_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/f_bench_$w.json 2>/dev/null || tail -3 gpurun_out/f_bench_$w.err
done
grep -o '"phases": {[^}]*}' gpurun_out/f_bench_c3.json
for w in c2 c1; do
  KEMR_MMA_PREFETCH=0 timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | python tools/benchsum.py "$w prefetch=0"
done
timeout 600 python tools/sweep_batch.py 43k > gpurun_out/f_sweep.jsonl 2> gpurun_out/f_sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/f_sweep.jsonl'):
    d=json.loads(l)
    print(d.get('shape','')[:4], 'B', d.get('B'), d.get('path','')[:8], 'scan_ms', d.get('scan_kernel_ms'), 'step_ms', d.get('step_ms'), 'frac', d.get('frac_of_measured_hbm'), 'unc', d.get('uncertified'), d.get('same_result_as_other_path'), d.get('error',''))
PY
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
for w in c2 c1; do
  KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-sharded > /dev/null 2> gpurun_out/f_dbg_$w.err
  grep "kemr mma dbg" gpurun_out/f_dbg_$w.err | head -2 | cut -c1-420 || tail -3 gpurun_out/f_dbg_$w.err
done
