#!/bin/bash
# Round 2, session S (one GPU): two lanes of the host-buffer index (kemr_index_share / submit / wait) -- parity, then the
# bench lines with the pipelined end-to-end figure beside the blocking one.
set -o pipefail
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_store.py -m gpu -q -x > $O/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/s_pytest.log
for w in c2 c1 c3 b64; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > $O/s_bench_$w.json 2> $O/s_bench_$w.err; echo "bench $w rc=$?"
  python tools/benchsum.py $w < $O/s_bench_$w.json 2>/dev/null || tail -5 $O/s_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('$O/s_bench_$w.json').read().strip().splitlines()[-1]); e=d['e2e']
    print('   e2e pipelined', round(e['value']), 'q/s', round(e['ms_per_step'],4), 'ms | blocking', round(e['blocking_call']['value']), round(e['blocking_call']['ms_per_step'],4), 'ms | device', round(d['ms_per_step'],4))
except Exception as ex: print('   parse failed', ex)
PY
done
