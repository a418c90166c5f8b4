#!/bin/bash
# Round 2, session L (one GPU): tcgen05 epilogue -- fold with four buffer entries in flight, per-query threshold shared by
# the lists of a query (KEMR_THR_SHARE=0 switches it off).  Parity, then c2 / c1 / b4096 / b64 with and without
# sharing, role cycles.
set -o pipefail
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/l_pytest_gpu.log
for share in 1 0; do
  for w in c2 c1 b4096 b64; do
    if [ $share = 0 ]; then export KEMR_THR_SHARE=0; else unset KEMR_THR_SHARE; fi
    timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > $O/l_bench_${w}_share$share.json 2> $O/l_bench_${w}_share$share.err; echo "bench $w share=$share rc=$?"
    python tools/benchsum.py $w share=$share < $O/l_bench_${w}_share$share.json 2>/dev/null || tail -3 $O/l_bench_${w}_share$share.err
  done
done
unset KEMR_THR_SHARE
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
if [ -f $DBG ]; then
  for w in c2 c1; do
    KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-sharded 2>&1 >/dev/null | grep "kemr mma dbg" | head -2 | tee $O/l_role_cycles_$w.txt | cut -c1-400
    KEMR_THR_SHARE=0 KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-sharded 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | tee $O/l_role_cycles_${w}_noshare.txt | cut -c1-400
  done
fi
timeout 600 python tools/sweep_batch.py 43k > $O/l_sweep.jsonl 2> $O/l_sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/l_sweep.jsonl'):
    d=json.loads(l)
    if 'tcgen05' in d.get('path',''):
        print(d.get('shape','')[:4], 'B', d.get('B'), d.get('path','')[:8], 'scan_ms', d.get('scan_kernel_ms'), 'step_ms', d.get('step_ms'), 'frac', d.get('frac_of_measured_hbm'), 'unc', d.get('uncertified'), d.get('same_result_as_other_path'), d.get('error',''))
PY
du -sh $O
