"""Synthetic refine scenes (TEST INFRASTRUCTURE) - SURVEY.md section 8(d).

A scene is everything ``Optimizer.optimize`` consumes for one detection: crop
intrinsics and size, the CSS NOCS prediction (3,128,128), a LIDAR crop and the
initial parameters.  The "prediction" is the oracle's own render of a ground-truth
pose/shape and the LIDAR points are sampled from the ground-truth surfels, so the
loop has something to converge to.  Everything is seeded.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import sdf_oracle as O


def intrinsics(size: int, focal_per_px: float = 2.25) -> torch.Tensor:
    """K = [[f,0,c],[0,f,c],[0,0,1]] with f = 144 at 64 px (cfg1) scaled with the crop."""
    f = focal_per_px * size
    c = size / 2.0
    return torch.tensor([[f, 0.0, c], [0.0, f, c], [0.0, 0.0, 1.0]], dtype=torch.float32)


def unit(v) -> np.ndarray:
    v = np.asarray(v, dtype=np.float32)
    return v / np.linalg.norm(v)


def make_scene(prior: O.DecoderParams, size: int = 64, density: int = 40, n_lidar: int = 400, seed: int = 1,
               width: Optional[int] = None, height: Optional[int] = None,
               gt=None, init=None, target_res: int = 128) -> Dict:
    """cfg1 (size=64) / cfg2 (size=256) of SURVEY.md 8(d); ``gt``/``init`` override the defaults."""
    rng = np.random.RandomState(seed)
    width = width or size
    height = height or size
    gt = gt or {"yaw": [0.6], "trans": [0.0, 0.0, 5.0], "scale": [2.0], "latent": unit([0.5, 0.7, 0.5])}
    init = init or {"yaw": [0.5], "trans": [0.1, 0.05, 5.0], "scale": [2.0], "latent": unit([0.6, 0.6, 0.5])}
    K = intrinsics(max(width, height))
    K[0, 2] = width / 2.0
    K[1, 2] = height / 2.0
    pts = O.lattice(density)
    with torch.no_grad():
        pass
    st = O.RefineState.create(gt["yaw"], gt["trans"], gt["scale"], gt["latent"])
    lat = torch.nn.functional.normalize(st.latent, dim=0)
    sdf, nrm, _ = O.sdf_and_normals(prior, lat, pts)
    sp, _, sn, _ = O.surface_points(pts, sdf.detach(), nrm)
    pose = O.yaw_pose(st.yaw, st.trans)
    # target rendered at the CSS resolution with the same field of view
    Kt = K.clone()
    Kt[0] *= target_res / width
    Kt[1] *= target_res / height
    with torch.no_grad():
        r = O.render(Kt, target_res, target_res, sp, sn, sn, pose, rot="dcm", output_nocs=True, tile_rows=8)
    nocs_pred = r["color"].numpy().astype(np.float32)
    front = r["xyzf"].numpy()
    if front.shape[0] > 0 and n_lidar > 0:
        pick = rng.choice(front.shape[0], size=n_lidar, replace=front.shape[0] < n_lidar)
        lidar = (front[pick] + rng.normal(0, 0.004, size=(n_lidar, 3))) * float(gt["scale"][0])
    else:
        lidar = np.zeros((0, 3))
    return {
        "K": K.numpy(), "crop_size": [height, width], "density": density,
        "nocs_pred": nocs_pred, "lidar": lidar.astype(np.float32),
        "init": {k: np.asarray(v, dtype=np.float32) for k, v in init.items()},
        "gt": {k: np.asarray(v, dtype=np.float32) for k, v in gt.items()},
        "weights": {"2d": 0.3, "3d": 0.5},
    }


def random_detection(prior: O.DecoderParams, seed: int, density: int = 40, target_res: int = 128) -> Dict:
    """cfg3-style detection: random latent, yaw, depth, crop size and LIDAR count."""
    rng = np.random.RandomState(1000 + seed)
    lat = unit(rng.normal(size=3) * 0.3 + np.array([0.5, 0.6, 0.5]))
    yaw = float(rng.uniform(-np.pi, np.pi))
    z = float(rng.uniform(4.0, 9.0))
    area = float(rng.uniform(32 ** 2, 64 ** 2))
    aspect = float(rng.uniform(0.6, 1.8))
    w = int(np.clip(round(np.sqrt(area * aspect)), 16, 96))
    h = int(np.clip(round(np.sqrt(area / aspect)), 16, 96))
    gt = {"yaw": [yaw], "trans": [float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.1, 0.1)), z],
          "scale": [2.0], "latent": lat}
    init = {"yaw": [yaw + float(rng.normal(0, 0.08))],
            "trans": [gt["trans"][0] + float(rng.normal(0, 0.05)), gt["trans"][1] + float(rng.normal(0, 0.03)),
                      z + float(rng.normal(0, 0.1))],
            "scale": [2.0 + float(rng.normal(0, 0.03))], "latent": unit(lat + rng.normal(0, 0.05, size=3))}
    sc = make_scene(prior, size=max(w, h), density=density, n_lidar=int(rng.randint(50, 800)), seed=seed,
                    width=w, height=h, gt=gt, init=init, target_res=target_res)
    # keep the object inside the crop: focal length so that the car spans ~80% of the crop at depth z
    return sc
