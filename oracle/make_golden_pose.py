"""Writes tests/golden/pose_ransac.npz by running the UNMODIFIED reference PoseEstimator
(TEST INFRASTRUCTURE; needs /root/reference, build container only).

    python -m oracle.make_golden_pose

Stores, per case, the seeded inputs, the numpy seed and what the reference's own
``PoseEstimator.init_pose_3d`` (utils/pose.py:84-233) returned.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import pose_oracle as PO
from . import ref_harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (scene kwargs, estimator type, numpy seed)
    "kabsch": (dict(seed=3), "kabsch", 11),
    "procrustes": (dict(seed=4, n_scene=220, n_outliers=90), "procrustes", 12),
    "kabsch_large": (dict(seed=5, n_model=3000, n_scene=700, n_outliers=150, yaw=-2.1, tra=(-2.0, 0.3, 14.0)), "kabsch", 13),
    "no_consensus": (dict(seed=6, n_scene=0, n_outliers=40), "kabsch", 14),
}


def main():
    ref_harness.load()                      # puts /root/reference on sys.path (open3d / pyquaternion stubbed)
    from utils.pose import PoseEstimator    # the reference's
    out = {}
    for name, (kw, typ, seed) in CASES.items():
        sc = PO.make_pose_scene(**kw)
        np.random.seed(seed)
        pose = PoseEstimator.init_pose_3d(sc['model_pts'].copy(), sc['model_cls'].copy(), sc['scene_pts'].copy(),
                                          sc['scene_cls'].copy(), type=typ, scale_model=sc['scale'])
        rng_after = np.random.randint(0, 2 ** 31 - 1)       # pins how much of the RNG stream was consumed
        for k in ('model_pts', 'model_cls', 'scene_pts', 'scene_cls'):
            out[f"{name}/{k}"] = sc[k]
        out[f"{name}/scale_model"] = np.float64(sc['scale'])
        out[f"{name}/seed"] = np.int64(seed)
        out[f"{name}/type"] = np.array(typ)
        out[f"{name}/found"] = np.bool_(pose is not None)
        out[f"{name}/rng_after"] = np.int64(rng_after)
        if pose is not None:
            out[f"{name}/scale"] = np.float64(pose['scale'])
            out[f"{name}/rot"] = np.asarray(pose['rot'], dtype=np.float64)
            out[f"{name}/tra"] = np.asarray(pose['tra'], dtype=np.float64)
            err = np.abs(np.asarray(pose['rot']) - sc['rot']).max()
            print(f"{name}: found, scale {float(pose['scale']):.4f}, |rot - gt| {err:.3e}, tra {np.asarray(pose['tra'])}")
        else:
            print(f"{name}: no pose")
    np.savez_compressed(os.path.join(GOLDEN, "pose_ransac.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "pose_ransac.npz"))


if __name__ == "__main__":
    sys.exit(main())
