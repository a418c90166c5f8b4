#!/bin/bash
# Round 2, session I: 16-warp streaming CTA + parallel ranking in the selection: phases, parity, bench c3 / c2, sweep.
set -o pipefail
mkdir -p gpurun_out
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
KEMR_LIB=$DBG timeout 300 python tools/select_phases.py 2>&1 | tee gpurun_out/i_select_phases.txt | tail -20
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/i_pytest_gpu.log
for w in c3 c2; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/i_bench_$w.json 2> gpurun_out/i_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/i_bench_$w.json 2>/dev/null || tail -3 gpurun_out/i_bench_$w.err
done
grep -o '"phases": {[^}]*}' gpurun_out/i_bench_c3.json
timeout 600 python tools/sweep_batch.py 43k > gpurun_out/i_sweep.jsonl 2> gpurun_out/i_sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/i_sweep.jsonl'):
    d=json.loads(l)
    print(d.get('shape','')[:4], 'B', d.get('B'), d.get('path','')[:8], 'scan_ms', d.get('scan_kernel_ms'), 'step_ms', d.get('step_ms'), 'frac', d.get('frac_of_measured_hbm'), 'unc', d.get('uncertified'), d.get('same_result_as_other_path'), d.get('error',''))
PY
