"""Synthetic frames for the cfg3 / cfg4 workloads of bench.py and the frame-loop tests.

KITTI is not available (no network), so a frame is 1-8 detections drawn from the committed
pool ``assets/bench_frames_pool.npz`` (written by oracle/make_bench_scene.py; this module only READS the file, it
does not import the oracle) with a seeded perturbation of the predicted latent and of the scene cloud, plus a rigid
LIDAR->camera matrix per frame.  ``make_frames(n, seed)`` is a pure function of its arguments, so every rank of a
multi-GPU run builds the same frame list.
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POOL = os.path.join(ROOT, "assets", "bench_frames_pool.npz")
_FIELDS = ("K", "crop_size", "nocs_pred", "lidar", "scene_pts", "scene_cls", "latent_pred", "orig_cam", "bbox")


def load_pool(path: str = POOL) -> List[Dict]:
    g = np.load(path)
    return [{k: g[f"p{i}_{k}"] for k in _FIELDS} for i in range(int(g["count"]))]


def make_frames(num_frames: int, seed: int = 0, pool: List[Dict] = None, min_det: int = 1, max_det: int = 8) -> List[Dict]:
    pool = pool if pool is not None else load_pool()
    frames = []
    for f in range(num_frames):
        rng = np.random.RandomState((seed * 7919 + f * 31 + 17) % (2 ** 31))
        a = rng.uniform(-0.2, 0.2)
        w2c = np.eye(4)
        w2c[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ \
            np.array([[0, -1, 0], [0, 0, -1], [1, 0, 0]], dtype=np.float64)
        w2c[:3, 3] = rng.uniform(-0.5, 0.5, size=3)
        dets = []
        for d in range(int(rng.randint(min_det, max_det + 1))):
            src = pool[int(rng.randint(len(pool)))]
            det = {k: src[k] for k in ("K", "nocs_pred", "lidar", "orig_cam")}
            det["crop_size"] = [int(v) for v in src["crop_size"]]
            det["bbox"] = [int(v) for v in src["bbox"]]
            lat = src["latent_pred"] + rng.normal(0, 0.02, size=src["latent_pred"].shape)
            det["latent_pred"] = (lat / np.linalg.norm(lat)).astype(np.float32)
            det["scene_pts"] = (src["scene_pts"] + rng.normal(0, 0.002, size=src["scene_pts"].shape)).astype(np.float32)
            det["scene_cls"] = src["scene_cls"]
            det["anno"] = {"name": "Car", "bbox": np.asarray(det["bbox"]), "alpha": 0.0, "dimensions": np.zeros(3),
                           "location": np.zeros(3), "rotation_y": 0.0, "score": 1.0}
            dets.append(det)
        frames.append({"detections": dets, "world_to_cam": w2c})
    return frames
