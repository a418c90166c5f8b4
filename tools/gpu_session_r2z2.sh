#!/bin/bash
# Two GPUs: the self-launching multi-GPU parity test on the final build, then the row-sharded metrics path over NCCL.
set -o pipefail
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multigpu.py -m gpu -x -q > gpurun_out/z2_pytest_multigpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z2_pytest_multigpu.log
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/time_sharded_metrics.py > gpurun_out/z2_sharded_metrics_n2.jsonl 2> gpurun_out/z2_sharded_metrics_n2.err; echo "sharded metrics rc=$?"
cat gpurun_out/z2_sharded_metrics_n2.jsonl; tail -3 gpurun_out/z2_sharded_metrics_n2.err
