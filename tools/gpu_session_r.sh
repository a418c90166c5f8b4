#!/bin/bash
for w in c2 c1; do for E in 1 2 1 2; do KEMR_MMA_EPI=$E timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline | python tools/benchsum.py "$w EPI=$E" | cut -c1-20,95-200; done; done
