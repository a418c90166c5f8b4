#!/bin/bash
# One GPU-box session: parity tests, every workload's bench line, ncu launch list + full capture of the scan kernel.
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for w in c2 c3 b64 c1 b4096; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < gpurun_out/bench_$w.json 2>/dev/null || tail -2 gpurun_out/bench_$w.err
done
CMD="python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:scan_mma -s 3 -c 2 -f -o gpurun_out/prof_scan_mma_c2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
