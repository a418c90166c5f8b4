#!/bin/bash
# quad-cluster (TMA multicast) scan: parity, then A/B against the pair kernel
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for w in c2 c1 b4096; do
for CL in 2 4; do
  KEMR_MMA_CL=$CL timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_cl$CL.json 2> gpurun_out/bench_${w}_cl$CL.err; echo "$w CL=$CL rc=$?"
  python tools/benchsum.py CL=$CL < gpurun_out/bench_${w}_cl$CL.json || tail -3 gpurun_out/bench_${w}_cl$CL.err
done; done
KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | tee gpurun_out/dbg_c2_quad.log
