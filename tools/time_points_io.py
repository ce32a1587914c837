"""Throughput of the native on-disk match writer / reader against numpy.savetxt / numpy.loadtxt loops on the bench workload's shape
(n pairs x ~2 544 matches, mkpts0 + mkpts1; default n = 512).  Host only."""
import os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pope_b200 import points_io

rng = np.random.default_rng(0)
n, cap = (int(sys.argv[1]) if len(sys.argv) > 1 else 512), 4800
counts = torch.from_numpy(rng.integers(2300, 2800, n).astype(np.int32))
out = {"mkpts0_f": torch.from_numpy(rng.uniform(0, 640, (n, cap, 2)).astype(np.float32)),
       "mkpts1_f": torch.from_numpy(rng.uniform(0, 640, (n, cap, 2)).astype(np.float32)), "counts": counts}
names = [f"{p:04d}" for p in range(n)]
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)      # RAM-backed: formatting, not the disk
for thr in (1, 0):
    dt = 1e9
    for rep in range(3):                                                           # best of 3
        shutil.rmtree(d + f"/b{thr}", ignore_errors=True)
        t0 = time.perf_counter()
        w = points_io.write_match_files(d + f"/b{thr}", names, out, threads=thr)
        dt = min(dt, time.perf_counter() - t0)
    print(f"native writer, threads={'all' if thr == 0 else thr}: {1e3 * dt:.1f} ms for {w} pairs ({int(counts.sum())} matches, "
          f"{n / dt:.0f} pairs/s)")
os.makedirs(d + "/np/mkpts0"); os.makedirs(d + "/np/mkpts1")
t0 = time.perf_counter()
n_np = min(n, 64)
for p in range(n_np):
    np.savetxt(d + f"/np/mkpts0/{names[p]}.txt", out["mkpts0_f"][p, :counts[p]].numpy())
    np.savetxt(d + f"/np/mkpts1/{names[p]}.txt", out["mkpts1_f"][p, :counts[p]].numpy())
dt = time.perf_counter() - t0
same = all(open(d + f"/np/{s}/{names[p]}.txt", "rb").read() == open(d + f"/b0/{s}/{names[p]}.txt", "rb").read()
           for p in range(n_np) for s in ("mkpts0", "mkpts1"))
print(f"numpy.savetxt loop over {n_np} pairs: {1e3 * dt:.1f} ms ({n_np / dt:.0f} pairs/s); files byte-identical: {same}; host cores: {os.cpu_count()}")
# the reader side (pose/dataset.py reads every file with numpy.loadtxt)
for thr in (1, 0):
    dt = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        back = points_io.read_match_files(d + "/b0", names, cap, threads=thr)
        dt = min(dt, time.perf_counter() - t0)
    print(f"native reader, threads={'all' if thr == 0 else thr}: {1e3 * dt:.1f} ms for {n} pairs ({n / dt:.0f} pairs/s)")
ok = all(torch.equal(back["mkpts0_f"][p, :counts[p]], out["mkpts0_f"][p, :counts[p]]) for p in range(n))
t0 = time.perf_counter()
for p in range(n_np):
    a = np.loadtxt(d + f"/b0/mkpts0/{names[p]}.txt", delimiter=" ")
    b = np.loadtxt(d + f"/b0/mkpts1/{names[p]}.txt", delimiter=" ")
dt = time.perf_counter() - t0
print(f"numpy.loadtxt loop over {n_np} pairs: {1e3 * dt:.1f} ms ({n_np / dt:.0f} pairs/s); round trip exact: {ok}")
# the two crops per pair (linemod.py:172-173 writes them with cv2.imwrite)
try:
    import cv2
    cv2.setNumThreads(1)
    yy, xx = np.mgrid[0:480, 0:640]
    img = np.stack([128 + 100 * np.sin(xx / 37.0 + yy / 91.0), 128 + 90 * np.cos(xx / 53.0), yy / 2.0], 2)
    img = (img + rng.normal(0, 4, img.shape)).clip(0, 255).astype(np.uint8)
    crops = [np.ascontiguousarray(img[: rng.integers(200, 480), : rng.integers(200, 640)]) for _ in range(128)]
    os.makedirs(d + "/png")
    t0 = time.perf_counter()
    for i, c in enumerate(crops):
        cv2.imwrite(d + f"/png/c{i}.png", c)
    t_cv = time.perf_counter() - t0
    line = f"128 crops (200-480 x 200-640 BGR): cv2.imwrite loop {1e3 * t_cv:.0f} ms"
    for thr in (1, 0):
        dt = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            points_io.imwrite_png_batch([d + f"/png/p{i}.png" for i in range(128)], crops, threads=thr)
            dt = min(dt, time.perf_counter() - t0)
        line += f"; native, threads={'all' if thr == 0 else thr}: {1e3 * dt:.0f} ms"
    same = all(np.array_equal(cv2.imread(d + f"/png/p{i}.png"), crops[i]) for i in range(128))
    size = lambda pre: sum(os.path.getsize(d + f"/png/{pre}{i}.png") for i in range(128)) >> 10
    print(line + f"; decoded equal: {same}; {size('c')} KB (cv2) vs {size('p')} KB")
except ImportError:
    print("cv2 not importable: PNG comparison skipped")
shutil.rmtree(d)
