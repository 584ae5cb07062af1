"""Pick the roofline-relevant counters out of `ncu --page raw --csv` (stdin)."""
import csv, sys
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
if len(rows) < 3:
    sys.exit("no rows")
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("----", d.get("Kernel Name", "?")[:80])
    for k in keys[1:]:
        if k in d:
            print(f"  {k:85s} {d[k]:>18s} {u.get(k,'')}")
