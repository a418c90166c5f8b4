#!/bin/bash
# A/B of two builds of the library on the SAME box: KEMR_LIB=<path> selects the library (libkemr_prev.so = the last
# commit's build, libkemr.so = the working tree's), alternating, three rounds.
set -o pipefail
mkdir -p gpurun_out
P=$PWD/knowledge_enhanced_multimodal_retrieval_b200
for round in 1 2 3; do
  for lib in libkemr_prev.so libkemr.so; do
    for w in ${WORKLOADS:-c1 c2}; do
      KEMR_LIB=$P/$lib timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded 2>/dev/null | python tools/benchsum.py $lib round$round
    done
  done
done
