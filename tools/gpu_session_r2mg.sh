#!/bin/bash
# Round 2, multi-GPU session: `gpurun --gpus N -- bash tools/gpu_session_r2mg.sh N`.  NCCL check of the row-sharded
# search (peer-memory exchange and NCCL all-gather against the single-GPU result), the bench line with its `sharded`
# record (10 M rows, strong scaling) at N ranks, the reference arm at N.
set -o pipefail
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multigpu.py -m gpu -q -x > gpurun_out/mg${N}_pytest.log 2>&1; echo "multigpu pytest rc=$?"; tail -4 gpurun_out/mg${N}_pytest.log
PORT=29533
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/mg${N}_bench.json').read().strip().splitlines()[-1])
    print('c2 N=${N}: value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']))
    for b in d['sharded']['batches']:
        print('  sharded B', b['queries_per_step'], 'rows/gpu', b['rows_per_gpu'], 'peer ms', round(b['peer']['ms_per_step'],4), 'nccl ms', round(b.get('nccl',{}).get('ms_per_step',0),4),
              'scan', round(b['peer']['scan_kernel_ms'],4), 'after(peer)', round(b['peer']['after_scan_ms'],4), 'after(nccl)', round(b.get('nccl',{}).get('after_scan_ms',0),4), 'frac', round(b['roofline']['frac'],3), 'equal', b.get('peer_equals_nccl'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/mg${N}_bench.err').read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/mg${N}_reference.json 2> gpurun_out/mg${N}_reference.err; echo "reference N=$N rc=$?"; cut -c1-300 gpurun_out/mg${N}_reference.json
