"""GPU bring-up check for the tcgen05 coarse kernels: compares row/column log-sum-exp and the match lists with the
fp32-FMA (SIMT) kernels on the same bf16 inputs, and with a torch fp32 reference of the log-sum-exp."""
import sys, os, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pope_b200 import _lib, ops, synth

dev = torch.device("cuda:0")
cases = [(1, (12, 16), (12, 16), 256), (2, (12, 16), (12, 16), 64), (3, (10, 14), (8, 9), 128), (2, (30, 40), (32, 32), 256),
         (1, (5, 5), (40, 50), 192), (2, (60, 80), (60, 80), 256)]
if len(sys.argv) > 1 and sys.argv[1] == "hires":      # BASELINE configs[3]: 960x1280 -> 19,200 coarse tokens per image
    cases = [(1, (120, 160), (120, 160), 256), (2, (120, 160), (120, 160), 256)]
elif len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
bad = 0
for n, hw0, hw1, C in cases:
    L, S = hw0[0] * hw0[1], hw1[0] * hw1[1]
    f0, f1 = synth.coarse_features(100 + L + C, n, L, S, C, sigma=0.9, dtype=torch.bfloat16)
    d0, d1 = f0.to(dev), f1.to(dev)
    out = {}
    for name, impl in (("simt", _lib.COARSE_SIMT), ("tc", _lib.COARSE_TCGEN05)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ops.coarse_match(d0, d1, hw0, hw1, 8.0, impl=impl)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        v = ops.scratch_views(r["workspace"], n, L, S)
        out[name] = (r, {k: v[k].clone() for k in v}, dt)
    # torch fp32 reference of the log-sum-exp in log2 units
    x = torch.einsum("nlc,nsc->nls", d0.float(), d1.float()) * (math.log2(math.e) / (C * 0.1))
    ref_r = torch.logsumexp(x * math.log(2), 2) / math.log(2)
    ref_c = torch.logsumexp(x * math.log(2), 1) / math.log(2)
    rs, vs, ts = out["simt"]; rt, vt, tt = out["tc"]
    e_simt = max((vs["lse_r"] - ref_r).abs().max().item(), (vs["lse_c"] - ref_c).abs().max().item())
    e_tc = max((vt["lse_r"] - ref_r).abs().max().item(), (vt["lse_c"] - ref_c).abs().max().item())
    ms, mt = rs.total(), rt.total()
    same = ms == mt and all(torch.equal(rs[k][:ms], rt[k][:mt]) for k in ("b_ids", "i_ids", "j_ids"))
    dconf = (rs["mconf"][:ms] - rt["mconf"][:mt]).abs().max().item() if same and ms else float("nan")
    ok = e_tc < 2e-3 and same
    bad += not ok
    print(f"n={n} L={L} S={S} C={C}: lse err simt {e_simt:.2e} tc {e_tc:.2e} | M simt {ms} tc {mt} same={same} "
          f"dconf={dconf:.2e} | t simt {ts*1e3:.2f} ms tc {tt*1e3:.2f} ms  {'OK' if ok else 'MISMATCH'}", flush=True)
print("tc_check:", "PASS" if bad == 0 else f"{bad} case(s) FAILED")
sys.exit(1 if bad else 0)
