#!/bin/bash
# after an MMA-issue change: smoke, parity tests, A/B of cluster sizes on the tensor-bound workloads, role cycles
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c2 c1 b4096}; do
for CL in ${CLS:-2 4}; do
  KEMR_MMA_CL=$CL timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_cl$CL.json 2> gpurun_out/bench_${w}_cl$CL.err; echo "$w CL=$CL rc=$?"
  python tools/benchsum.py CL=$CL < gpurun_out/bench_${w}_cl$CL.json || tail -3 gpurun_out/bench_${w}_cl$CL.err
done; done
for CL in ${CLS:-2 4}; do
KEMR_MMA_CL=$CL KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | sed 's/.*stages/stages/'
done
