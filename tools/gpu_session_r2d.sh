#!/bin/bash
# Round 2, session D: parity; bench c2 / c1 / c3 with the mask epilogue (variant 3) and, for comparison, variant 2;
# batch sweep with the streaming kernel v2 and the rebalanced single-CTA tcgen05 plan; role cycles.
set -o pipefail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/d_pytest_gpu.log
for w in c2 c1 c3; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > gpurun_out/d_bench_$w.json 2> gpurun_out/d_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/d_bench_$w.json 2>/dev/null || tail -3 gpurun_out/d_bench_$w.err
done
grep -o '"phases": {[^}]*}' gpurun_out/d_bench_c3.json
for w in c2 c1; do
  KEMR_MMA_EPI=2 timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | python tools/benchsum.py "$w epi=2"
done
timeout 600 python tools/sweep_batch.py > gpurun_out/d_sweep.jsonl 2> gpurun_out/d_sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/d_sweep.jsonl'):
    d=json.loads(l)
    print(d.get('shape','')[:4], 'B', d.get('B'), d.get('path','')[:8], 'scan_ms', d.get('scan_kernel_ms'), 'step_ms', d.get('step_ms'), 'frac', d.get('frac_of_measured_hbm'), 'unc', d.get('uncertified'), d.get('same_result_as_other_path'), d.get('error',''))
PY
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
for w in c2 c1; do
  KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-sharded 2>&1 >/dev/null | grep "kemr mma dbg" | head -2 | tee gpurun_out/d_dbg_$w.log | cut -c1-420
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench_default.json 2> gpurun_out/d_bench_default.err; echo "bench default rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/d_bench_default.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    for b in d['sharded']['batches']:
        print('sharded B', b['queries_per_step'], 'ms', b['ms_per_step'], 'scan', b['peer']['scan_kernel_ms'], 'after', b['peer']['after_scan_ms'], 'frac', b['roofline']['frac'], b['roofline']['bound'])
except Exception as e:
    print('default parse failed', e); print(open('gpurun_out/d_bench_default.err').read()[-1500:])
PY
