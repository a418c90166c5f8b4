"""Print the metrics of an `ncu --page raw --csv` export that matter for the scan kernel (diagnostics helper)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2:]
keys = sys.argv[2:] or ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                        'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor', 'lts__t_bytes.sum', 'lts__t_sectors.sum',
                        'lts__throughput', 'sm__throughput', 'gpu__dram_throughput', 'launch__registers_per_thread',
                        'sm__warps_active', 'sm__cycles_elapsed.max', 'lts__t_sector_hit_rate', 'l1tex__m_xbar2l1tex',
                        'smsp__inst_executed.sum', 'launch__grid_size', 'sm__cycles_active.avg', 'tma', 'lts__t_sectors_srcunit_tex']
for i, h in enumerate(hdr):
    if any(k in h for k in keys):
        print(h, '[' + units[i] + ']', [v[i] for v in vals])
