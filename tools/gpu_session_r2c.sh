#!/bin/bash
# Round 2, session C: parity (stripe-synchronised tcgen05 plan, peer exchange, prepared search), bench c2/c1/c3,
# role cycles at 43k rows for small batches, ncu --set full with source of the C1 scan and the streaming kernel.
set -o pipefail
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/c_pytest_gpu.log
for w in c2 c1 c3; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_$w.json 2> gpurun_out/c_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/c_bench_$w.json 2>/dev/null || tail -3 gpurun_out/c_bench_$w.err
done
DBG=$PWD/knowledge_enhanced_multimodal_retrieval_b200/libkemr_debug.so
for w in c2 c1; do
  KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -2 | tee gpurun_out/c_dbg_$w.log | cut -c1-420
done
KEMR_LIB=$DBG KEMR_MMA_DEBUG=1 timeout 300 python tools/sweep_batch.py 43k 2> gpurun_out/c_sweep_dbg.err > gpurun_out/c_sweep_dbg.jsonl; echo "sweep dbg rc=$?"
grep "kemr mma dbg" gpurun_out/c_sweep_dbg.err | head -8 | cut -c1-420
CMD="python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"scan_mma" -s 4 -c 1 -f -o gpurun_out/prof_c1_scan $CMD > gpurun_out/c_ncu_c1.log 2>&1
echo "ncu c1 rc=$?"
CMD="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/c_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"scan_stream" -s 4 -c 1 -f -o gpurun_out/prof_c3_stream $CMD > gpurun_out/c_ncu_c3.log 2>&1
echo "ncu c3 rc=$?"
ls -la gpurun_out/*.ncu-rep
