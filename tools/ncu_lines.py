"""Top CUDA source lines by warp-stall samples from `ncu -i X --page source --print-source cuda,sass --csv` (diagnostics)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = []
fname = ''
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        si = hdr.index('Warp Stall Sampling (All Samples)')
        ei = hdr.index('Instructions Executed')
        continue
    if hdr and len(r) > max(si, ei) and r[0].isdigit() and r[si].isdigit():   # a CUDA source line with its totals
        out.append((int(r[si]), int(r[ei] or 0), fname, r[0], r[1][:110]))
tot = sum(o[0] for o in out) or 1
print('total samples', tot)
for o in sorted(out, reverse=True)[:top]:
    print(f'{100 * o[0] / tot:5.1f}% inst {o[1]:10d} {o[2]}:{o[3]}  {o[4]}')
