#!/bin/bash
mkdir -p gpurun_out
for w in b4096 b64; do
for E in 0 1; do
  KEMR_MMA_EPI=$E timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline | python tools/benchsum.py "$w EPI=$E"
done; done
for w in c2 c1; do timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline | python tools/benchsum.py "$w auto"; done
