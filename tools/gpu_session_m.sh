#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_$name.json 2> gpurun_out/bench_${W}_$name.err
  python tools/benchsum.py $name < gpurun_out/bench_${W}_$name.json || tail -3 gpurun_out/bench_${W}_$name.err
}
W=c2; run cl2_ds KEMR_MMA_CL=2; run cl2_nods KEMR_MMA_CL=2 KEMR_MMA_NO_DS=1; run cl4_ds KEMR_MMA_CL=4
W=c1; run cl2 KEMR_MMA_CL=2; run cl4 KEMR_MMA_CL=4
KEMR_MMA_CL=2 KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | sed 's/.*stages/stages/'
