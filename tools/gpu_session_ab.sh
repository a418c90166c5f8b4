#!/bin/bash
# A/B on the SAME box, alternating, three rounds: either two builds of the library (LIBS="libkemr_prev.so libkemr.so",
# selected with KEMR_LIB) or one build under different environment switches (ENVS="KEMR_SEL_W=4 KEMR_SEL_W=2").
# Boxes differ by up to ~7 % (C2 scan 87.9 us on one, 94.9 us on another, same build), so only same-box pairs compare.
set -o pipefail
mkdir -p gpurun_out
P=$PWD/knowledge_enhanced_multimodal_retrieval_b200
for round in 1 2 3; do
  if [ -n "$ENVS" ]; then
    for e in $ENVS; do
      for w in ${WORKLOADS:-c1 c2}; do
        env $e timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | python tools/benchsum.py $e round$round
      done
    done
  else
    for lib in ${LIBS:-libkemr_prev.so libkemr.so}; do
      for w in ${WORKLOADS:-c1 c2}; do
        KEMR_LIB=$P/$lib timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | python tools/benchsum.py $lib round$round
      done
    done
  fi
done
