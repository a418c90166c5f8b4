#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload b4096 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_b4096.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows=list(csv.reader(l for l in open('gpurun_out/launches_b4096.csv') if l.startswith('"')))
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size'); bi=h.index('Block Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault((r[ki][:70], r[gi], r[bi]),[]).append(float(r[vi].replace(',','')))
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1), 'us')
PY
