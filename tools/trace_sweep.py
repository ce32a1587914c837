"""Per-tile timeline of the tcgen05 sweeps (developer diagnostics): POPE_TC_TRACE=<0|2> python tools/trace_sweep.py"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pope_b200 import _lib, ops, synth
mode = os.environ.get("POPE_TC_TRACE", "0")
n = 64
dev = torch.device("cuda:0")
f0, f1 = synth.coarse_features(1234, n, 4800, 4800, 256, dtype=torch.bfloat16)
d0, d1 = f0.to(dev), f1.to(dev)
ws = torch.empty(_lib.lib().pope_coarse_workspace_bytes(n, 4800, 4800), dtype=torch.uint8, device=dev)
for _ in range(3):
    r = ops.coarse_match(d0, d1, (60, 80), (60, 80), 8.0, workspace=ws)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (8 * 2 * 512))()
got = _lib.lib().pope_debug_trace_read(buf, len(buf))
a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64).reshape(2, 512, 8)
mma, epi = a[0], a[1]
ntile = int((mma[:, 0] > 0).sum())
print("mode", mode, "records", got, "tiles traced", ntile)
mma, epi = mma[:ntile], epi[:ntile]
first = mma[:, 0] < 0             # top bit of the first stamp marks the first tile of a unit
mma[:, 0] &= (1 << 63) - 1
t0 = mma[0, 0]
w_acc = mma[:, 1] - mma[:, 0]     # issuer waits for the epilogue to free the accumulator stage
w_b = mma[:, 2] - mma[:, 1]       # ... then for the first operand chunk
issue = mma[:, 3] - mma[:, 2]     # ... then issues the tile's MMAs (incl. waits for later chunks)
period = np.diff(mma[:, 3])
e_wait = epi[:, 1] - epi[:, 0]    # epilogue warp waits for the accumulator
e_hand = epi[:, 2] - epi[:, 1]    # accumulator ready -> stage handed back
e_tot = epi[:, 3] - epi[:, 1]     # accumulator ready -> tile done
def st(x, m=None):
    x = x if m is None else x[m]
    return "mean %7.0f  p50 %7.0f  p90 %7.0f  max %7.0f" % (x.mean(), np.median(x), np.percentile(x, 90), x.max())
print("issuer: wait acc_empty      ", st(w_acc))
print("issuer: wait first b_full   ", st(w_b), "| first tile of unit:", st(w_b, first))
print("issuer: issue tile          ", st(issue))
print("issuer: tile period         ", st(period), "| across unit boundary:", st(period, first[1:]))
print("epilogue: wait acc_full     ", st(e_wait), "| first tile of unit:", st(e_wait, first))
print("epilogue: ready -> hand-back", st(e_hand))
print("epilogue: ready -> done     ", st(e_tot))
ch = np.diff(np.concatenate([mma[:, 2:3], mma[:, 4:8]], 1), axis=1)   # per K-chunk: wait for its operands + issue 4 MMAs + commit
print("issuer: per-chunk (wait + 4 MMA issues + commit), chunks 0..3:", " | ".join("%4.0f" % x for x in np.median(ch, 0)))
print("total cycles %d for %d tiles = %.0f per tile" % (mma[-1, 3] - t0, ntile, (mma[-1, 3] - t0) / ntile))
