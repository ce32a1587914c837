"""ncu --csv launch log (several metrics per launch) -> one line per launch.  usage: python tools/ncu_launch_table.py log.csv [max_rows]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[iid], r[ik][:44]), {})[r[im]] = r[iv]
short = {"gpu__time_duration.sum": "t_ns", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
         "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%"}
for n, ((i, k), m) in enumerate(d.items()):
    if len(sys.argv) > 2 and n >= int(sys.argv[2]):
        break
    print(i, k, " ".join(f"{short.get(kk, kk)}={vv}" for kk, vv in m.items()))
