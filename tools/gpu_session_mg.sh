#!/bin/bash
# multi-GPU session: N = number of GPUs of the box ($1)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
timeout 600 $TR tests/run_multigpu_check.py 2>&1 | tail -3
for w in ${WORKLOADS:-c2 c4 c5_b64 c5_b8192}; do
  timeout 900 $TR bench.py --gpus $N --workload $w --steps ${STEPS:-5} --warmup 3 $EXTRA > gpurun_out/mg${N}_$w.json 2> gpurun_out/mg${N}_$w.err; echo "bench $w N=$N rc=$?"
  python tools/benchsum.py "$w@$N" < gpurun_out/mg${N}_$w.json 2>/dev/null || tail -5 gpurun_out/mg${N}_$w.err
done
