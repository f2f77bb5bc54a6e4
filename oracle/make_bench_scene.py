"""Writes the synthetic inputs bench.py feeds to both arms (TEST INFRASTRUCTURE; run once in the build
container: python -m oracle.make_bench_scene [cfg2] [frames]):

  assets/bench_scene_256.npz   cfg2: one 256x256 detection
  assets/bench_frames_pool.npz cfg3 / cfg4: a pool of ragged detections (crop <= 96 px, 50-800 LIDAR points) with
                               everything the frame loop of refine_css.py:94-250 consumes after the CSS network:
                               NOCS prediction, LIDAR crop, crop intrinsics, full-image camera + 2D box, the
                               NOCS-coloured scene cloud of the pose initialisation and the predicted latent.
                               bench.py draws its frames (1-8 detections each) from this pool.
"""
import os
import sys

import numpy as np
import torch

from . import prior as P
from . import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def frames_pool(prior, count=12, density=24):
    from . import sdf_oracle as O
    out = {"count": count}
    for i in range(count):
        sc = scenes.random_detection(prior, 100 + i, density=density)
        rng = np.random.RandomState(7000 + i)
        gt = sc["gt"]
        # the scene cloud of the pose initialisation (refine_css.py:157-161): camera-frame points with the
        # NOCS colours the CSS network predicts for them = front-facing ground-truth surfels, metric scale
        st = O.RefineState.create(gt["yaw"], gt["trans"], gt["scale"], gt["latent"])
        lat = torch.nn.functional.normalize(st.latent, dim=0)
        pts = O.lattice(density)
        sdf, nrm, _ = O.sdf_and_normals(prior, lat, pts)
        sp, _, sn, _ = O.surface_points(pts, sdf.detach(), nrm)
        v, m, c, front = O.to_camera(sp, sn, sn, O.yaw_pose(st.yaw, st.trans), "dcm", True)
        v_all = v.numpy().astype(np.float64)
        v, col = v[front].numpy(), ((c[front] + 1) / 2).numpy()
        n_scene = int(rng.randint(250, 700))
        pick = rng.choice(v.shape[0], size=n_scene, replace=v.shape[0] < n_scene)
        scene_pts = (v[pick] + rng.normal(0, 0.004, size=(n_scene, 3))) * float(gt["scale"][0])
        scene_cls = np.clip(col[pick] + rng.normal(0, 0.01, size=(n_scene, 3)), 0, 1)
        # 5 % outliers: background points with arbitrary colours
        n_out = n_scene // 20
        scene_pts[:n_out] += rng.normal(0, 0.6, size=(n_out, 3))
        h, w = sc["crop_size"]
        left, top = int(rng.randint(0, 1100)), int(rng.randint(0, 280))
        orig_cam = sc["K"].astype(np.float64).copy()
        orig_cam[0, 2] += left
        orig_cam[1, 2] += top
        # the 2D detection box: tight around the object in the full image (what a detector returns; the height
        # re-estimation of refine_css.py:176-186 compares it with the projected initial model)
        px = v_all @ orig_cam.T
        px = px[:, :2] / px[:, 2:3]
        bbox = np.asarray([np.floor(px[:, 0].min()), np.floor(px[:, 1].min()), np.ceil(px[:, 0].max()),
                           np.ceil(px[:, 1].max())], dtype=np.int64)
        fields = {"K": sc["K"], "crop_size": np.asarray(sc["crop_size"]), "nocs_pred": sc["nocs_pred"],
                  "lidar": sc["lidar"], "scene_pts": scene_pts.astype(np.float32),
                  "scene_cls": scene_cls.astype(np.float32), "latent_pred": sc["init"]["latent"],
                  "orig_cam": orig_cam, "bbox": bbox}
        for k, val in sc["gt"].items():
            fields["gt_" + k] = val
        for k, val in fields.items():
            out[f"p{i}_{k}"] = val
        print(f"pool {i}: crop {w}x{h}, {sc['lidar'].shape[0]} lidar, {n_scene} scene points", file=sys.stderr)
    path = os.path.join(ROOT, "assets", "bench_frames_pool.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


def main():
    prior = P.load_prior(os.path.join(ROOT, "assets", "deepsdf_synth.pt"))
    what = sys.argv[1:] or ["cfg2", "frames"]
    if "frames" in what:
        frames_pool(prior)
    if "cfg2" not in what:
        return
    sc = scenes.make_scene(prior, size=256, density=40)
    out = {"K": sc["K"], "crop_size": np.asarray(sc["crop_size"]), "density": 40, "nocs_pred": sc["nocs_pred"],
           "lidar": sc["lidar"], "w2d": sc["weights"]["2d"], "w3d": sc["weights"]["3d"]}
    for k, v in sc["init"].items():
        out["init_" + k] = v
    for k, v in sc["gt"].items():
        out["gt_" + k] = v
    path = os.path.join(ROOT, "assets", "bench_scene_256.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
