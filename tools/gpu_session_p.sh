#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload b4096 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 3 -c 1 -f -o gpurun_out/prof_select_b4096 $CMD > gpurun_out/ncu_full_sel.log 2>&1
echo "ncu full rc=$?"
