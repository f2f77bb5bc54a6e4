"""Second batch of golden files written by the UNMODIFIED reference (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_r2 [traj50] [raster] [latent256] [kitti_label] [refine_ragged]

SURVEY.md Appendix C sizes the first batch (oracle/make_golden.py) left out:
  * T8  - a 50-step ``Optimizer.optimize`` trajectory at 64x64 / D=40 (cfg1), parameters and
          the printed losses after every step (the Adam state persists across the 50 calls of
          ``optimize(1)`` exactly as it does inside one call of ``optimize(50)``:
          pipelines/optimizer.py:46-52 builds the solver once per Optimizer);
  * T5  - ``Rasterer`` maps, point lists and gradients at 64x64 and at 256x256 (the reference
          evaluated in 16-row pixel tiles, SURVEY.md 8(c): bit-exact with the untiled call);
  * T2  - a latent-256 decoder;
  * a14 - ``get_kitti_label`` (utils/refinement.py:501-562 with roty_in_bev / alpha_in_bev);
  * cfg3 - ``Optimizer.optimize`` on three ragged detections (non-square crops, different LIDAR
          counts), 6 steps each, for the batched engine.
Needs /root/reference: build container only.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import re
import sys

import numpy as np
import torch

from . import make_golden as G1
from . import prior as P
from . import ref_harness
from . import scenes
from . import sdf_oracle as O

GOLDEN = G1.GOLDEN
STOCK_PRIOR = G1.STOCK_PRIOR
_np = G1._np

_LOSS_RE = re.compile(r"ITER (\d+) \| Losses: 2D - (\S+), 3D - (\S+), Total - (\S+)")


def _run_optimize_steps(ref, dec, sc, steps, density):
    """``steps`` calls of the reference's optimize(1): parameters after every call and the losses it printed
    ([w2d*loss2d, w3d*loss3d, total]; NaN rows for skipped iterations)."""
    params = {k: v.copy() for k, v in sc["init"].items()}
    opt = ref.Optimizer(params, torch.device("cpu"), sc["weights"])
    ref.grid_module.grads.clear()
    grid = ref.Grid3D(density)
    traj, losses = [], []
    for it in range(steps):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            opt.optimize(1, torch.tensor(sc["nocs_pred"]), sc["lidar"], dec, grid, torch.tensor(sc["K"]),
                         sc["crop_size"], viz_type=None)
        m = _LOSS_RE.search(buf.getvalue())
        losses.append([float(m.group(2)), float(m.group(3)), float(m.group(4))] if m else [np.nan] * 3)
        traj.append(np.concatenate([_np(params[k]).reshape(-1) for k in ("yaw", "trans", "scale", "latent")]))
        print(f"  step {it + 1}/{steps}: {losses[-1]}", file=sys.stderr, flush=True)
    return np.stack(traj), np.asarray(losses, dtype=np.float64)


def _scene_fields(sc, density):
    out = {"K": sc["K"], "crop_size": np.asarray(sc["crop_size"]), "density": density, "nocs_pred": sc["nocs_pred"],
           "lidar": sc["lidar"], "w2d": sc["weights"]["2d"], "w3d": sc["weights"]["3d"]}
    for k, v in sc["init"].items():
        out["init_" + k] = v
    return out


def golden_traj50(ref):
    dec, L = ref.setup_dsdf(STOCK_PRIOR, precision=torch.float32)
    prior = P.load_prior(STOCK_PRIOR)
    sc = scenes.make_scene(prior, size=64, density=40)          # cfg1 of SURVEY.md 8(d)
    traj, losses = _run_optimize_steps(ref, dec, sc, 50, 40)
    np.savez_compressed(os.path.join(GOLDEN, "refine_traj50_64.npz"), traj=traj, losses=losses, **_scene_fields(sc, 40))


def golden_refine_ragged(ref):
    """Three cfg3-style detections (random crop shapes / LIDAR counts), 6 reference steps each at D=20."""
    dec, L = ref.setup_dsdf(STOCK_PRIOR, precision=torch.float32)
    prior = P.load_prior(STOCK_PRIOR)
    out = {}
    for i, seed in enumerate((0, 1, 2)):
        sc = scenes.random_detection(prior, seed, density=20)
        traj, losses = _run_optimize_steps(ref, dec, sc, 6, 20)
        for k, v in _scene_fields(sc, 20).items():
            out[f"d{i}_{k}"] = v
        out[f"d{i}_traj"], out[f"d{i}_losses"] = traj, losses
    np.savez_compressed(os.path.join(GOLDEN, "refine_ragged.npz"), count=3, **out)


def pattern_cotangent(shape, salt):
    """Deterministic pseudo-random cotangent in [0, 1) (a multiplicative hash of the flat index): the
    tests recompute it, so the 256x256 golden does not carry 2 MB of incompressible noise."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    h = (idx * np.uint64(2654435761) + np.uint64(salt * 97 + 13)) % np.uint64(1 << 20)
    return torch.from_numpy((h.astype(np.float64) / float(1 << 20)).astype(np.float32).reshape(shape))


def _tiled_raster(ref, K, w, h, coords, normals, pose, tile_rows, gen):
    """Reference Rasterer evaluated tile by tile (rows [r0, r0+tile_rows)): every raster op is per pixel, so
    replacing ``renderer.grid`` by a slice of the pixel grid and ``res_y_px`` by the slice height gives the
    rows of the full call bit for bit.  Gradients accumulate over the tiles."""
    ras = ref.Rasterer(K, (w, h))
    full_grid = ras.grid.clone()
    maps = {k: [] for k in ("color", "mask", "depth", "normals")}
    cots = {}
    for i, (k, c) in enumerate((("color", 3), ("mask", 1), ("depth", 1), ("normals", 3))):
        cots[k] = pattern_cotangent((c, h, w), i)
    g_acc = [torch.zeros_like(coords), torch.zeros_like(normals), torch.zeros_like(pose)]
    points = None
    for r0 in range(0, h, tile_rows):
        r1 = min(h, r0 + tile_rows)
        ras.grid = full_grid[:, r0 * w:r1 * w]
        ras.res_y_px = r1 - r0
        c = coords.detach().clone().requires_grad_(True)
        n = normals.detach().clone().requires_grad_(True)
        p = pose.detach().clone().requires_grad_(True)
        rendering, pts = ras(c, n, n, p, rot="dcm", primitives="disc", bg=None, output_depth=True,
                             output_normals=True, output_nocs=True, output_mask=True, output_points=True)
        scalar = sum((rendering[k] * cots[k][:, r0:r1]).sum() for k in maps)
        if r0 == 0:
            points = {k: _np(v) for k, v in pts.items()}
            cots["xyzf"] = pattern_cotangent(tuple(pts["xyzf"].shape), 7)
            scalar = scalar + (pts["xyzf"] * cots["xyzf"]).sum()
        gs = torch.autograd.grad(scalar, [c, n, p])
        for a, g in zip(g_acc, gs):
            a += g
        for k in maps:
            maps[k].append(_np(rendering[k]))
    rendering = {k: np.concatenate(v, axis=1) for k, v in maps.items()}
    return rendering, points, cots, g_acc


def golden_raster_large(ref):
    """T5 at 64x64 (one call and 16-row tiles: asserted identical) and 256x256 (tiled)."""
    prior = P.load_prior(STOCK_PRIOR)
    sp, sn = G1._surfels(prior, 24)
    pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.05, -0.02, 4.0])).detach()
    for size, tile in ((64, 16), (256, 16)):
        gen = torch.Generator().manual_seed(100 + size)
        K = scenes.intrinsics(size)
        rendering, points, cots, g_acc = _tiled_raster(ref, K, size, size, sp, sn, pose, tile, gen)
        if size == 64:      # the tiling trick itself, checked against the plain call
            gen2 = torch.Generator().manual_seed(100 + size)
            r_full, p_full, _, g_full = _tiled_raster(ref, K, size, size, sp, sn, pose, size, gen2)
            for k in rendering:
                assert np.array_equal(rendering[k], r_full[k]), k
            for a, b in zip(g_acc, g_full):
                assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), "tiled gradient differs"
        out = {"K": _np(K), "width": size, "height": size, "coords": _np(sp), "normals": _np(sn), "pose": _np(pose),
               "g_coords": _np(g_acc[0]), "g_normals": _np(g_acc[1]), "g_pose": _np(g_acc[2])}
        for k, v in rendering.items():
            out["r_" + k] = v
        # the cotangents are pattern_cotangent(shape, i): recomputed by the tests, not stored
        for k, v in points.items():
            out["p_" + k] = v
        np.savez_compressed(os.path.join(GOLDEN, f"raster_dcm_{size}x{size}.npz"), **out)


def golden_latent256(ref):
    spec = O.DecoderSpec(256, [288, 288, 288], latent_in=(2,), norm_layers=(0, 1, 2), weight_norm=True)
    gen = torch.Generator().manual_seed(17)
    sd = P.random_prior(spec, seed=13)
    dec = G1._ref_decoder(ref, spec, sd)
    n = 200
    lat = torch.nn.functional.normalize(torch.randn(n, spec.latent_size, generator=gen), dim=1)
    xyz = torch.rand(n, 3, generator=gen) * 2 - 1
    inp = torch.cat([lat, xyz], 1).requires_grad_(True)
    sdf, _ = dec(inp)
    (g,) = torch.autograd.grad(sdf.sum(), inp)
    out = {"inputs": _np(inp), "sdf": _np(sdf), "dinput": _np(g),
           "spec_json": np.frombuffer(json.dumps(spec.to_json()).encode(), dtype=np.uint8)}
    for k, v in sd.items():
        out["sd::" + k] = _np(v)
    np.savez_compressed(os.path.join(GOLDEN, "decoder_latent256.npz"), **out)


def golden_kitti_label(ref):
    """get_kitti_label of the reference for a handful of refined states and LIDAR->camera matrices."""
    import utils.refinement as rtools
    dec, L = ref.setup_dsdf(STOCK_PRIOR, precision=torch.float32)
    rng = np.random.RandomState(21)
    cases = []
    for i in range(6):
        latent = torch.tensor(rng.normal(size=3) * 0.3 + np.array([0.5, 0.6, 0.5]), dtype=torch.float32)
        scale = torch.tensor([float(rng.uniform(1.6, 2.4))])
        trans = torch.tensor([float(rng.uniform(-2, 2)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(3, 12))])
        yaw = torch.tensor([float(rng.uniform(-np.pi, np.pi))])
        # a rigid LIDAR->camera transform (KITTI's Tr_velo_to_cam has this form), identity for case 0
        p_WC = np.eye(4)
        if i:
            a = rng.uniform(-0.2, 0.2)
            p_WC[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) @ \
                np.array([[0, -1, 0], [0, 0, -1], [1, 0, 0]], dtype=np.float64)
            p_WC[:3, 3] = rng.uniform(-0.5, 0.5, size=3)
        bbox = [10.0 * i, 5.0, 80.0 + i, 60.0]
        ref.grid_module.grads.clear()
        grid = ref.Grid3D(30)
        label, pts, cam_T = rtools.get_kitti_label(dec, grid, latent, scale, trans, yaw, p_WC, bbox)
        cases.append({"latent": _np(latent), "scale": _np(scale), "trans": _np(trans), "yaw": _np(yaw), "p_WC": p_WC,
                      "bbox": np.asarray(bbox), "location": np.asarray(label["location"], dtype=np.float64),
                      "dimensions": np.asarray(label["dimensions"], dtype=np.float64),
                      "rotation_y": np.float64(label["rotation_y"]), "alpha": np.float64(label["alpha"]),
                      "score": np.float64(label["score"]), "cam_T": cam_T, "n_points": pts.shape[0],
                      "points_min": pts.min(0), "points_max": pts.max(0)})
    out = {"count": len(cases)}
    for i, c in enumerate(cases):
        for k, v in c.items():
            out[f"c{i}_{k}"] = v
    np.savez_compressed(os.path.join(GOLDEN, "kitti_label.npz"), **out)


TASKS = {"traj50": golden_traj50, "raster": golden_raster_large, "latent256": golden_latent256,
         "kitti_label": golden_kitti_label, "refine_ragged": golden_refine_ragged}


def main():
    if not ref_harness.available():
        sys.exit("reference tree not found - golden files can only be generated in the build container")
    torch.manual_seed(1)
    np.random.seed(1)
    ref = ref_harness.load()
    for name in (sys.argv[1:] or list(TASKS)):
        print("==", name, file=sys.stderr, flush=True)
        TASKS[name](ref)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
