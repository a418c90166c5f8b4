#!/bin/bash
# quick: smoke + short A/B + role cycles + one full ncu capture of the scan kernel (c2)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in ${WORKLOADS:-c2 b4096}; do
for CL in ${CLS:-2 4}; do
  KEMR_MMA_CL=$CL timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_cl$CL.json 2> gpurun_out/bench_${w}_cl$CL.err; echo "$w CL=$CL rc=$?"
  python tools/benchsum.py CL=$CL < gpurun_out/bench_${w}_cl$CL.json || tail -3 gpurun_out/bench_${w}_cl$CL.err
done; done
KEMR_MMA_CL=2 KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | sed 's/.*stages/stages/'
CMD="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline"
KEMR_MMA_CL=2 $CMD > gpurun_out/plain.log 2>&1 &&
KEMR_MMA_CL=2 ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 3 -c 1 -f -o gpurun_out/prof_scan_mma_c2_v7 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
