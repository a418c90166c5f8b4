#!/bin/bash
# Round 2, session N (one GPU): final evidence of the round: the round's evidence run after the container was re-created.  Parity suite, the bench
# line (with its `sharded` record and cpu_baseline), the reference arm, every single-GPU workload, the batch sweeps,
# role cycles, ncu launch lists (c2, c3) and `--set full` captures of the tcgen05 scan, the streaming search kernel
# and the selection kernel.  Everything lands in gpurun_out/n_*; summaries are copied to profiles/r02_* by hand.
set -o pipefail
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -1
if [ -n "$WITH_PYTEST" ]; then timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/n_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $O/n_pytest_gpu.log; fi
timeout 300 python tools/time_host_search.py > $O/n_host_search.jsonl 2> $O/n_host_search.err; echo "host search rc=$?"; cat $O/n_host_search.jsonl; tail -3 $O/n_host_search.err
timeout 900 python bench.py > $O/n_bench_default.json 2> $O/n_bench_default.err; echo "bench default rc=$?"
python tools/benchsum.py default < $O/n_bench_default.json || tail -5 $O/n_bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/n_bench_reference.json 2> $O/n_bench_reference.err; echo "bench reference rc=$?"; cut -c1-260 $O/n_bench_reference.json
for w in c3 c1 b64 b4096; do
  extra="--no-cpu-baseline"; [ $w = c3 ] && extra=""
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 $extra > $O/n_bench_$w.json 2> $O/n_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < $O/n_bench_$w.json 2>/dev/null || tail -3 $O/n_bench_$w.err
done
grep -o '"phases": {[^}]*}' $O/n_bench_c3.json
timeout 900 python tools/sweep_batch.py > $O/n_sweep.jsonl 2> $O/n_sweep.err; echo "sweep rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/n_sweep.jsonl'):
    d=json.loads(l)
    print(d.get('shape','')[:4], 'B', d.get('B'), d.get('path','')[:8], 'scan_ms', d.get('scan_kernel_ms'), 'step_ms', d.get('step_ms'), 'frac', d.get('frac_of_measured_hbm'), 'unc', d.get('uncertified'), d.get('same_result_as_other_path'), d.get('error',''))
PY
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
if [ -f $DBG ]; then
  KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline --no-sharded 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | tee $O/n_role_cycles_c2.txt | cut -c1-120
  KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline --no-sharded 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | tee $O/n_role_cycles_c1.txt | cut -c1-120
  KEMR_LIB=$DBG timeout 300 python tools/select_phases.py > $O/n_select_phases.txt 2>&1; tail -12 $O/n_select_phases.txt
fi
# ---- ncu: launch lists (only after the same command exited 0 without ncu), then full captures
for w in c2 c3; do
  CMD="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-sharded"
  $CMD > $O/n_plain_$w.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/n_launches_$w.csv $CMD > $O/n_ncu_list_$w.log 2>&1
  echo "ncu list $w rc=$?"
done
CMD="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded"
timeout 600 ncu --set full --clock-control none -k regex:"scan_mma|select_kernel" -s 6 -c 2 -f -o $O/n_prof_c2 $CMD > $O/n_ncu_full_c2.log 2>&1; echo "ncu full c2 rc=$?"
CMD="python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded"
timeout 600 ncu --set full --clock-control none -k regex:"scan_stream" -s 4 -c 1 -f -o $O/n_prof_c3 $CMD > $O/n_ncu_full_c3.log 2>&1; echo "ncu full c3 rc=$?"
CMD="python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded"
timeout 600 ncu --set full --clock-control none -k regex:"scan_mma" -s 4 -c 1 -f -o $O/n_prof_c1 $CMD > $O/n_ncu_full_c1.log 2>&1; echo "ncu full c1 rc=$?"
CMD="python bench.py --workload b64 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded"
timeout 600 ncu --set full --clock-control none -k regex:"scan_mma" -s 4 -c 1 -f -o $O/n_prof_b64 $CMD > $O/n_ncu_full_b64.log 2>&1; echo "ncu full b64 rc=$?"
for p in c2 c3 c1 b64; do
  if [ -f $O/n_prof_$p.ncu-rep ]; then
    ncu -i $O/n_prof_$p.ncu-rep --page raw --csv > $O/n_prof_${p}_raw.csv 2>/dev/null
    ncu -i $O/n_prof_$p.ncu-rep --page details > $O/n_prof_${p}_details.txt 2>/dev/null
  fi
done
rm -f $O/n_prof_*.ncu-rep      # summaries travel back, the reports stay on the box (gpurun_out/ is capped at 64 MiB)
du -sh $O
ls -la $O | head -60
