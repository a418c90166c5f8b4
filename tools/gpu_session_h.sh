#!/bin/bash
mkdir -p gpurun_out
./tools/microbench/fp64_rates | tee gpurun_out/fp64_rates.log
CMD="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline"
KEMR_SELECT_WARPS=2 $CMD > gpurun_out/plain.log 2>&1 &&
KEMR_SELECT_WARPS=2 ncu --set full --clock-control none --import-source on -k regex:select_query -s 3 -c 1 -f -o gpurun_out/prof_select_query_c2 $CMD > gpurun_out/ncu_full_sel.log 2>&1
echo "ncu full rc=$?"
