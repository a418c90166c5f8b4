#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/diag_pair_dense.py 2>&1 | tail -20
CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 3 -c 1 -f -o gpurun_out/prof_scan_mma_pair_c2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
