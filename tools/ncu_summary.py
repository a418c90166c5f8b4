"""Summary of an `ncu --set full` capture for profiles/: the metrics the roofline argument rests on, per captured kernel,
from `ncu -i X.ncu-rep --page raw --csv`.

    python tools/ncu_summary.py gpurun_out/j_prof_c2_raw.csv > profiles/r02_ncu_full_c2.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum.per_second",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tma.sum", "smsp__inst_executed_op_tma_ld.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for v in vals:
        print("kernel:", v[col["Kernel Name"]], "| grid", v[col.get("Grid Size", 0)], "| block", v[col.get("Block Size", 0)])
        for k in KEYS:
            hit = [h for h in hdr if h == k] or [h for h in hdr if h.startswith(k)]
            for h in hit[:3]:
                print(f"  {h:88s} {v[col[h]]:>18s} {units[col[h]]}")
        print()


if __name__ == "__main__":
    main()
