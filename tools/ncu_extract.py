"""ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) -> compact per-kernel table of the metrics we cite."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum"]
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
w.writerow(["kernel"] + [f"{k} [{units[h.index(k)]}]" if k in h else k for k in KEYS])
for r in rows[2:]:
    w.writerow([r[h.index("Kernel Name")][:90]] + [r[h.index(k)] if k in h else "" for k in KEYS])
