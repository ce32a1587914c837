"""Developer diagnostics for the lazily shifted single sweep: runs a large-norm case, prints the flags, the per-pair flags and
where the row / column log-sum-exp differ from the oracle.   python tools/debug_shift.py [sigma]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import _lib, ops, synth
sig = float(sys.argv[1]) if len(sys.argv) > 1 else 49.0 / 16.0
h0, w0, h1, w1 = 40, 48, 36, 56
L, S = h0 * w0, h1 * w1
dev = torch.device("cuda:0")
f0, f1 = synth.coarse_features(91, 2, L, S, 256, sigma=sig, dtype=torch.bfloat16)
r = ops.coarse_match(f0.to(dev), f1.to(dev), (h0, w0), (h1, w1), 8.0, impl=_lib.COARSE_TCGEN05)
torch.cuda.synchronize()
v = ops.scratch_views(r["workspace"], 2, L, S)
print("flags", r.flags(), "pairflag", v["pairflag"].tolist(), "M", r.total())
scale = 1.4426950408889634 / (256 * 0.1)
X = torch.einsum("nlc,nsc->nls", f0.double(), f1.double()) * scale
lr, lc = torch.logsumexp(X * 0.6931471805599453, 2) / 0.6931471805599453, torch.logsumexp(X * 0.6931471805599453, 1) / 0.6931471805599453
dr = (v["lse_r"].cpu().double() - lr).abs()
dc = (v["lse_c"].cpu().double() - lc).abs()
print("x range", float(X.min()), float(X.max()))
print("lse_r max err", float(dr.max()), "nonfinite", int((~torch.isfinite(v['lse_r'])).sum()), "worst rows", torch.topk(dr.flatten(), 5).indices.tolist())
print("lse_c max err", float(dc.max()), "nonfinite", int((~torch.isfinite(v['lse_c'])).sum()), "worst cols", torch.topk(dc.flatten(), 5).indices.tolist())
cs = v["cshift"].cpu()
print("cshift min/max", float(cs.min()), float(cs.max()), "nonfinite", int((~torch.isfinite(cs)).sum()))
cnt = v["cand_cnt"].cpu().to(torch.int32) & 0xffff
nib = torch.stack([(cnt >> (4 * k)) & 0xf for k in range(4)], -1)
print("max list fill", int(nib.max()), "lists with >= 6", int((nib >= 6).sum()))
