#!/bin/bash
# Round 2, session M (one GPU): selection kernel with a register cap (KEMR_SEL_MINB = 0 / 6 / 8 CTAs per SM), C1 with
# short lists + virtual parts + shared threshold (KEMR_MMA_KLIST=8 KEMR_THR_SHARE=1).
set -o pipefail
mkdir -p gpurun_out
O=gpurun_out
for mb in 0 6 8; do
  for w in c2 c1; do
    KEMR_SEL_MINB=$mb timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > $O/m_bench_${w}_minb$mb.json 2> $O/m_bench_${w}_minb$mb.err; echo "bench $w minb=$mb rc=$?"
    python tools/benchsum.py $w minb=$mb < $O/m_bench_${w}_minb$mb.json 2>/dev/null || tail -3 $O/m_bench_${w}_minb$mb.err
  done
done
for kl in "8 1" "8 0" "16 1"; do
  set -- $kl
  KEMR_MMA_KLIST=$1 KEMR_THR_SHARE=$2 timeout 400 python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline --no-sharded > $O/m_bench_c1_k$1_s$2.json 2> $O/m_bench_c1_k$1_s$2.err; echo "bench c1 K=$1 share=$2 rc=$?"
  python tools/benchsum.py c1 K=$1 share=$2 < $O/m_bench_c1_k$1_s$2.json 2>/dev/null || tail -3 $O/m_bench_c1_k$1_s$2.err
done
KEMR_SEL_MINB=8 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/m_pytest_minb8.log 2>&1; echo "pytest minb8 rc=$?"; tail -3 $O/m_pytest_minb8.log
