#!/bin/bash
# Round 2, session K (one GPU): pipelined host-buffer search (query chunks over the copy engine), request in the kernel
# parameters for one / two queries, KG hits pre-scored by the scanning CTAs, programmatic dependent launch of the
# selection kernel, tensor-map cache.  Parity first, then the host-search breakdown and the bench lines.
set -o pipefail
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > $O/k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/k_pytest_gpu.log
timeout 300 python tools/time_host_search.py > $O/k_host_search.jsonl 2> $O/k_host_search.err; echo "host search rc=$?"; cat $O/k_host_search.jsonl; tail -3 $O/k_host_search.err
for ch in "0" "512" "256,256" "256"; do
  echo "KEMR_E2E_CHUNKS=$ch"; KEMR_E2E_CHUNKS=$ch timeout 300 python tools/time_host_search.py c2 2>&1 | grep -v pageable | cut -c1-220
done
echo "KEMR_NO_PDL=1"; KEMR_NO_PDL=1 timeout 300 python tools/time_host_search.py c2 2>&1 | grep -v pageable | cut -c1-220
echo "KEMR_NO_INLINE_REQUEST=1"; KEMR_NO_INLINE_REQUEST=1 timeout 300 python tools/time_host_search.py c3 2>&1 | grep -v pageable | cut -c1-220
timeout 900 python bench.py > $O/k_bench_default.json 2> $O/k_bench_default.err; echo "bench default rc=$?"
python tools/benchsum.py default < $O/k_bench_default.json || tail -5 $O/k_bench_default.err
for w in c3 c1; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/k_bench_$w.json 2> $O/k_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < $O/k_bench_$w.json 2>/dev/null || tail -3 $O/k_bench_$w.err
done
grep -o '"phases": {[^}]*}' $O/k_bench_c3.json
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
[ -f $DBG ] && { KEMR_LIB=$DBG timeout 300 python tools/select_phases.py > $O/k_select_phases.txt 2>&1; tail -16 $O/k_select_phases.txt; }
timeout 600 python tools/sweep_batch.py 43k > $O/k_sweep.jsonl 2> $O/k_sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/k_sweep.jsonl'):
    d=json.loads(l)
    print(d.get('shape','')[:4], 'B', d.get('B'), d.get('path','')[:8], 'scan_ms', d.get('scan_kernel_ms'), 'step_ms', d.get('step_ms'), 'frac', d.get('frac_of_measured_hbm'), 'unc', d.get('uncertified'), d.get('same_result_as_other_path'), d.get('error',''))
PY
du -sh $O
