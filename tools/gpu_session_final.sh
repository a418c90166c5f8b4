#!/bin/bash
# Round evidence: parity tests, default bench line (+reference arm), every workload, role cycles,
# ncu launch list + full captures of the scan and select kernels.
set -o pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
python tools/benchsum.py default < gpurun_out/bench_default.json || tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
for w in c3 c1 b64 b4096; do
  timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/bench_$w.json 2>/dev/null || tail -3 gpurun_out/bench_$w.err
done
KEMR_MMA_DEBUG=1 timeout 300 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "kemr mma dbg" | head -1 | tee gpurun_out/dbg_c2.log | cut -c1-80
CMD="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"scan_mma|select_kernel" -s 6 -c 2 -f -o gpurun_out/prof_c2_final $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
