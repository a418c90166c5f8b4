#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_index or retrieval_engine or quantize" 2>&1 | tail -3
for w in c2 c3; do
for z in 0 1; do
  if [ $z = 1 ]; then export KEMR_NO_ZERO_COPY=1; else unset KEMR_NO_ZERO_COPY; fi
  timeout 300 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${w}_nozc$z.json 2> gpurun_out/bench_${w}_nozc$z.err
  python tools/benchsum.py "$w no_zero_copy=$z" < gpurun_out/bench_${w}_nozc$z.json || tail -3 gpurun_out/bench_${w}_nozc$z.err
done; done
