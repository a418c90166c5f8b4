#!/bin/bash
# Round 2, session Q (one GPU): validation of the final build -- parity suite, bench line, c1 / c3 / b4096, smoke().
set -o pipefail
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/q_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/q_bench_default.json 2> $O/q_bench_default.err; echo "bench default rc=$?"
python tools/benchsum.py default < $O/q_bench_default.json || tail -5 $O/q_bench_default.err
for w in c1 c3 b4096; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/q_bench_$w.json 2> $O/q_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < $O/q_bench_$w.json 2>/dev/null || tail -3 $O/q_bench_$w.err
done
CMD="python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded"
$CMD > $O/q_plain_c1.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/q_launches_c1.csv $CMD > $O/q_ncu_list_c1.log 2>&1; echo "ncu list c1 rc=$?"
