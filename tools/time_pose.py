"""Times pope_estimate_pose_batch (CUDA events, match lists resident on the device) against the reference's per-pair CPU
path (cv2.findEssentialMat + cv2.recoverPose as in src/utils/metrics.py:69-94, restated inline so that the tool does not need
/root/reference) on the same synthetic scenes.   python tools/time_pose.py [pairs] [matches] [outlier_frac]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle.gen_golden_pose import scene
from pope_b200 import pose

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
outl = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
rng = np.random.default_rng(0)
sc = [scene(rng, m, outl, 0.1, 600.0, 650.0) for _ in range(n)]
dev = torch.device("cuda:0")
mk0 = torch.from_numpy(np.concatenate([s[0] for s in sc])).to(dev)
mk1 = torch.from_numpy(np.concatenate([s[1] for s in sc])).to(dev)
counts = torch.tensor([len(s[0]) for s in sc], dtype=torch.int32, device=dev)
K0, K1 = torch.from_numpy(np.stack([s[2] for s in sc])).to(dev), torch.from_numpy(np.stack([s[3] for s in sc])).to(dev)


def ang(a, b):
    return float(np.degrees(np.arccos(np.clip((np.trace(a.T @ b) - 1) / 2, -1, 1))))


for conf in (0.99, 0.99999):
    ws = None
    for _ in range(3):
        out = pose.estimate_pose_batch(mk0, mk1, counts, K0, K1, 0.5, conf, workspace=ws)
        ws = out["workspace"]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        out = pose.estimate_pose_batch(mk0, mk1, counts, K0, K1, 0.5, conf, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # the call is stream-ordered and free of host synchronisation, so it can be captured once and replayed as a CUDA graph
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gout = pose.estimate_pose_batch(mk0, mk1, counts, K0, K1, 0.5, conf, workspace=ws)
    graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / reps
    same = all(torch.equal(gout[k], out[k]) for k in ("R", "t", "inliers", "iters"))
    R = out["R"].cpu().numpy()
    errs = [ang(R[p], sc[p][4]) for p in range(n)]
    line = (f"conf {conf}: {n} pairs x {m} matches ({outl:.0%} outliers): {ms:.3f} ms per batch = {n / ms * 1e3:,.0f} pairs/s "
            f"(as a replayed CUDA graph {ms_graph:.3f} ms, same result: {same}); "
            f"iters median {int(out['iters'].median())} max {int(out['iters'].max())}; R error median {np.median(errs):.3f} max {max(errs):.3f} deg; "
            f"inlier frac {float(out['n_inliers'].float().mean()) / m:.3f}")
    try:
        import cv2
        cv2.setNumThreads(1)
        k = min(n, 8)
        t0 = time.perf_counter()
        cerr = []
        for p in range(k):
            p0, p1, k0, k1 = sc[p][0], sc[p][1], sc[p][2], sc[p][3]
            a = (p0 - k0[[0, 1], [2, 2]][None]) / k0[[0, 1], [0, 1]][None]
            b = (p1 - k1[[0, 1], [2, 2]][None]) / k1[[0, 1], [0, 1]][None]
            thr = 0.5 / np.mean([k0[0, 0], k1[1, 1], k0[0, 0], k1[1, 1]])
            E, mask = cv2.findEssentialMat(a, b, np.eye(3), threshold=thr, prob=conf, method=cv2.RANSAC)
            _, Rc, tc, _ = cv2.recoverPose(E[:3], a, b, np.eye(3), 1e9, mask=mask)
            cerr.append(ang(Rc, sc[p][4]))
        cpu_ms = (time.perf_counter() - t0) / k * 1e3
        line += f" | cv2 (1 thread, {k} pairs): {cpu_ms:.1f} ms per pair = {1e3 / cpu_ms:.1f} pairs/s, R error median {np.median(cerr):.3f} deg"
    except ImportError:
        line += " | cv2 not importable here"
    print(line, flush=True)
