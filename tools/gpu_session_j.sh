#!/bin/bash
# ingress experiment on the pair kernel (merged mode, c2): per-role cycles with query / gallery loads skipped after the first tile
for SK in 0 1 2 3; do
  echo "SKIP=$SK"
  KEMR_MMA_CL=2 KEMR_MMA_DEBUG_SKIP=$SK KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep "kemr mma dbg" | head -1 | sed 's/.*producer/producer/'
done
