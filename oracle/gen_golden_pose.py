"""Generates tests/golden/pose_cv2.npz: synthetic two-view scenes and what the UNMODIFIED reference `estimate_pose`
(/root/reference/src/utils/metrics.py:69-94, i.e. opencv-python's findEssentialMat + recoverPose) returns for them.
Run in the build container (needs /root/reference and cv2); the fixture travels, this script's inputs do not.

    python oracle/gen_golden_pose.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("POPE_REFERENCE_ROOT", "/root/reference")


def load_reference_metrics():
    """metrics.py imports loguru and kornia at module level (metrics.py:5-7); estimate_pose uses neither, so empty
    stand-ins are enough to import the file unmodified."""
    for name in ("loguru", "kornia", "kornia.geometry", "kornia.geometry.epipolar", "kornia.geometry.conversions"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.modules["loguru"].__dict__.setdefault("logger", None)
    sys.modules["kornia.geometry.epipolar"].__dict__.setdefault("numeric", None)
    sys.modules["kornia.geometry.conversions"].__dict__.setdefault("convert_points_to_homogeneous", None)
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "src", "utils", "metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def scene(rng, m, outlier_frac, noise, f0, f1, kind="general"):
    """m correspondences of a random rigid motion seen by two pinhole cameras; the first outlier_frac are random.
    kind: "general" (points in a 3-D box), "planar" (all points on one slanted plane: a degenerate configuration for
    uncalibrated solvers, which the five-point solver must still handle), "small_baseline" (translation 1/25 of the
    others: the essential matrix is close to the pure-rotation degeneracy)."""
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = rng.uniform(0.1, 0.6)
    kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * kx + (1 - np.cos(ang)) * kx @ kx
    t = rng.normal(size=3)
    t /= np.linalg.norm(t)
    X = np.stack([rng.uniform(-1, 1, m), rng.uniform(-1, 1, m), rng.uniform(3, 6, m)], 1)
    if kind == "planar":
        X[:, 2] = 4.5 + 0.4 * X[:, 0] - 0.3 * X[:, 1]
    K0 = np.array([[f0, 0, 320.0], [0, f0 * 1.02, 240.0], [0, 0, 1]])
    K1 = np.array([[f1, 0, 300.0], [0, f1 * 0.98, 260.0], [0, 0, 1]])
    p0 = X @ K0.T
    p0 = p0[:, :2] / p0[:, 2:]
    X1 = X @ R.T + (0.02 if kind == "small_baseline" else 0.5) * t
    p1 = X1 @ K1.T
    p1 = p1[:, :2] / p1[:, 2:]
    p0 = p0 + rng.normal(size=p0.shape) * noise
    p1 = p1 + rng.normal(size=p1.shape) * noise
    no = int(m * outlier_frac)
    p1[:no] = rng.uniform(0, 600, (no, 2))
    planted = np.arange(m) >= no                      # the correspondences that follow the motion (up to the pixel noise)
    return p0.astype(np.float32), p1.astype(np.float32), K0, K1, R, t, planted


def main():
    ref = load_reference_metrics()
    rng = np.random.default_rng(20260)
    spec = [(400, 0.1, 0.1), (1200, 0.3, 0.1), (2500, 0.5, 0.1), (60, 0.2, 0.05), (3, 0.0, 0.1), (0, 0.0, 0.1),
            (800, 0.4, 0.2), (5, 0.0, 0.0), (1500, 0.2, 0.1), (300, 0.6, 0.1),
            # appended in round 2 (the earlier scenes keep their random stream): near-degenerate geometry
            (600, 0.3, 0.1, "planar"), (600, 0.3, 0.05, "small_baseline"), (1000, 0.2, 0.1, "planar")]
    thresh = 0.5
    confs = {"hi": 0.99999, "lo": 0.99}       # metrics.py:69 default / eval_onepose_json.py:164
    scenes = [scene(rng, sp[0], sp[1], sp[2], rng.uniform(500, 700), rng.uniform(500, 700), *sp[3:]) for sp in spec]
    arrays = dict(mkpts0=np.concatenate([s[0] for s in scenes]), mkpts1=np.concatenate([s[1] for s in scenes]),
                  counts=np.array([s[0] for s in spec], dtype=np.int32), K0=np.stack([s[2] for s in scenes]),
                  K1=np.stack([s[3] for s in scenes]), R_gt=np.stack([s[4] for s in scenes]),
                  t_gt=np.stack([s[5] for s in scenes]), planted=np.concatenate([s[6] for s in scenes]),
                  thresh=np.float64(thresh))
    for tag, conf in confs.items():
        Rc, tc, st, masks = [], [], [], []
        for (p0, p1, K0, K1, _, _, _) in scenes:
            ret = ref.estimate_pose(p0, p1, K0, K1, thresh, conf)
            if ret is None:
                st.append(0); Rc.append(np.zeros((3, 3))); tc.append(np.zeros(3)); masks.append(np.zeros(len(p0), dtype=bool))
            else:
                st.append(1); Rc.append(ret[0]); tc.append(ret[1]); masks.append(ret[2])
            print(f"conf {conf} pair {len(st) - 1}: m={len(p0)} status={st[-1]} inliers={int(masks[-1].sum())}")
        arrays.update({f"conf_{tag}": np.float64(conf), f"R_cv2_{tag}": np.stack(Rc), f"t_cv2_{tag}": np.stack(tc),
                       f"status_cv2_{tag}": np.array(st, dtype=np.int32), f"inliers_cv2_{tag}": np.concatenate(masks)})
    out = os.path.join(ROOT, "tests", "golden", "pose_cv2.npz")
    np.savez_compressed(out, **arrays)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
