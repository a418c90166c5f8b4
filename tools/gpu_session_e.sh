#!/bin/bash
mkdir -p gpurun_out
python tools/diag_pair_count.py > gpurun_out/diag_count.log 2>&1; grep "PAIR=" gpurun_out/diag_count.log
for p in 1 0; do
KEMR_MMA_PAIR=$p KEMR_MMA_DEBUG=1 KEMR_MMA_DEBUG_TRACE=1 timeout 300 python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma" > gpurun_out/trace_pair$p.log
done
tail -100 gpurun_out/trace_pair1.log | head -60
