"""Summarise an `ncu --page source --csv` export: opcode mix (executed warp-instructions), stall reasons, hottest lines.
usage: python tools/ncu_src_summary.py src.csv [kernel_index] [warp_chunks]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
ki = int(sys.argv[2]) if len(sys.argv) > 2 else 0
seg = rows[starts[ki] + 2: starts[ki + 1] if ki + 1 < len(starts) else None]
h = rows[starts[ki] + 1]
print("kernel:", rows[starts[ki]][1][:100], "(%d kernels in file)" % len(starts))
ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot, samp = collections.Counter(), collections.Counter()
lines = []
for r in seg:
    if len(r) <= ie or not r[ie].isdigit():
        continue
    src = r[ia]
    op = (src.split()[1] if src.strip().startswith("@") else src.split()[0]).split(".")[0]
    tot[op] += int(r[ie])
    samp[op] += int(r[isamp])
    lines.append((int(r[isamp]), int(r[ie]), r[0][-5:], src[:90]))
T, S = sum(tot.values()), sum(samp.values())
unit = float(sys.argv[3]) if len(sys.argv) > 3 else None
print("warp-instructions", T, "samples", S)
for op, n in tot.most_common(22):
    print(f"  {op:10s} {n:12d} {100 * n / T:5.1f}%  samples {100 * samp[op] / S:5.1f}%" + (f"  per-unit {n / unit:7.1f}" if unit else ""))
stalls = collections.Counter()
for r in seg:
    for i, c in enumerate(h):
        if c.startswith("stall_") and "Not Issued" not in c and i < len(r) and r[i].isdigit():
            stalls[c] += int(r[i])
SS = sum(stalls.values())
print("stalls:", ", ".join(f"{k[6:]} {100 * v / SS:.1f}%" for k, v in stalls.most_common(9)))
print("hottest lines:")
for l in sorted(lines, reverse=True)[:14]:
    print("  ", l)
