"""Writes tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

    python -m oracle.make_golden            # needs /root/reference (build container only)

The reference ships no golden vectors for this path, so
these files are the pin: each one stores the seeded inputs and what the
reference's own code (Grid3D, Decoder, Rasterer, Optimizer) returned for them.
``tests/test_oracle_golden.py`` checks the oracle restatement against them;
the GPU parity tests check the CUDA path against them and against the oracle.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import torch

from . import prior as P
from . import ref_harness
from . import scenes
from . import sdf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
STOCK_PRIOR = os.path.join(ROOT, "assets", "deepsdf_synth.pt")

SMALL_SPECS = {
    "wn_skip": O.DecoderSpec(3, [64, 64, 64, 64], latent_in=(2,), norm_layers=(0, 1, 2, 3), weight_norm=True),
    "layernorm": O.DecoderSpec(3, [48, 40, 48], latent_in=(2,), norm_layers=(0, 1, 2), weight_norm=False),
    "xyz_in_all": O.DecoderSpec(3, [64, 64, 64], latent_in=(), norm_layers=(0, 1, 2), weight_norm=True,
                                xyz_in_all=True),
    "use_tanh": O.DecoderSpec(3, [32, 32], latent_in=(), norm_layers=(), weight_norm=False, use_tanh=True),
    "latent8": O.DecoderSpec(8, [96, 96, 96, 96], latent_in=(3,), norm_layers=(0, 1, 2, 3), weight_norm=True),
}


def _np(t):
    return t.detach().cpu().numpy()


def _ref_decoder(ref, spec: O.DecoderSpec, sd):
    """Loads a checkpoint through the reference's own setup_dsdf."""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "net.pt")
        P.save_prior(path, spec, sd)
        dec, L = ref.setup_dsdf(path, precision=torch.float32)
    return dec


def golden_lattice(ref):
    out = {}
    for d in (8, 9, 12):
        out[f"points_{d}"] = _np(ref.Grid3D(d).points)
    for d in (30, 40, 41):
        out[f"sha256_{d}"] = np.frombuffer(
            hashlib.sha256(_np(ref.Grid3D(d).points).tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "lattice.npz"), **out)


def golden_decoders(ref):
    gen = torch.Generator().manual_seed(7)
    for name, spec in SMALL_SPECS.items():
        sd = P.random_prior(spec, seed=11)
        dec = _ref_decoder(ref, spec, sd)
        n = 300
        lat = torch.nn.functional.normalize(torch.randn(n, spec.latent_size, generator=gen), dim=1)
        xyz = torch.rand(n, 3, generator=gen) * 2 - 1
        inp = torch.cat([lat, xyz], 1).requires_grad_(True)
        sdf, _ = dec(inp)
        (g,) = torch.autograd.grad(sdf.sum(), inp)
        out = {"inputs": _np(inp), "sdf": _np(sdf), "dinput": _np(g),
               "spec_json": np.frombuffer(json.dumps(spec.to_json()).encode(), dtype=np.uint8)}
        for k, v in sd.items():
            out["sd::" + k] = _np(v)
        np.savez_compressed(os.path.join(GOLDEN, f"decoder_{name}.npz"), **out)


def golden_stock(ref):
    """Stock 8x512 prior: sdf / gradient on a lattice subsample, surface extraction at D=16."""
    dec, L = ref.setup_dsdf(STOCK_PRIOR, precision=torch.float32)
    lat = torch.nn.functional.normalize(torch.tensor([0.5, 0.7, 0.5]), dim=0)
    ref.grid_module.grads.clear()
    grid = ref.Grid3D(16)
    inp = torch.cat([lat.expand(grid.points.shape[0], -1), grid.points], 1)
    sdf, _ = dec(inp)
    pts, nocs, nrm = grid.get_surface_points(sdf)
    keep = (sdf.abs() < 0.03).squeeze(1)
    np.savez_compressed(
        os.path.join(GOLDEN, "stock_surface_d16.npz"), latent_unit=_np(lat), sdf=_np(sdf),
        raw_grad_normalised=_np(ref.grid_module.grads['grid_points']), keep=_np(keep),
        surf_pts=_np(pts), surf_nocs=_np(nocs), surf_nrm=_np(nrm))
    return dec


def _surfels(dec_params: O.DecoderParams, density=24):
    pts = O.lattice(density)
    lat = torch.nn.functional.normalize(torch.tensor([0.5, 0.7, 0.5]), dim=0)
    sdf, nrm, _ = O.sdf_and_normals(dec_params, lat, pts)
    sp, _, sn, _ = O.surface_points(pts, sdf.detach(), nrm)
    return sp.detach(), sn.detach()


def golden_raster(ref):
    prior = P.load_prior(STOCK_PRIOR)
    sp, sn = _surfels(prior, 24)
    gen = torch.Generator().manual_seed(3)
    for tag, (w, h), rot in (("dcm_45x22", (45, 22), "dcm"), ("dcm_32x32", (32, 32), "dcm"),
                             ("quat_40x30", (40, 30), "quat")):
        K = scenes.intrinsics(max(w, h))
        K[0, 2], K[1, 2] = w / 2.0, h / 2.0
        coords = sp.clone().requires_grad_(True)
        normals = sn.clone().requires_grad_(True)
        if rot == "dcm":
            pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.05, -0.02, 4.0])).detach().requires_grad_(True)
        else:
            q = torch.tensor([0.92, 0.05, 0.38, -0.04])
            pose = torch.cat([q, torch.tensor([0.05, -0.02, 4.0])]).requires_grad_(True)
        ras = ref.Rasterer(K, (w, h))
        res = ras(coords, normals, normals, pose, rot=rot, primitives='disc', bg=None, output_depth=True,
                  output_normals=True, output_nocs=True, output_mask=True, output_points=(rot == "dcm"))
        rendering, points = res if rot == "dcm" else (res, None)
        # fixed random cotangents -> scalar -> gradients
        cot = {k: torch.rand(v.shape, generator=gen) for k, v in rendering.items()}
        scalar = sum((rendering[k] * cot[k]).sum() for k in rendering)
        if points is not None:
            cot["xyzf"] = torch.rand(points["xyzf"].shape, generator=gen)
            scalar = scalar + (points["xyzf"] * cot["xyzf"]).sum()
        g_coords, g_normals, g_pose = torch.autograd.grad(scalar, [coords, normals, pose])
        out = {"K": _np(K), "width": w, "height": h, "coords": _np(coords), "normals": _np(normals),
               "pose": _np(pose), "g_coords": _np(g_coords), "g_normals": _np(g_normals), "g_pose": _np(g_pose)}
        for k, v in rendering.items():
            out["r_" + k] = _np(v)
            out["cot_" + k] = _np(cot[k])
        if points is not None:
            for k, v in points.items():
                out["p_" + k] = _np(v)
            out["cot_xyzf"] = _np(cot["xyzf"])
        np.savez_compressed(os.path.join(GOLDEN, f"raster_{tag}.npz"), **out)


def golden_primitives(ref):
    """The two circle primitives and background compositing (rasterer.py:93-126) on one small view: colour and
    mask maps and the gradient with respect to the points, per (primitive, background) combination."""
    prior = P.load_prior(STOCK_PRIOR)
    sp, sn = _surfels(prior, 24)
    w, h = 40, 30
    K = scenes.intrinsics(max(w, h))
    K[0, 2], K[1, 2] = w / 2.0, h / 2.0
    pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.05, -0.02, 4.0])).detach()
    bg = torch.rand(3, h, w, generator=torch.Generator().manual_seed(11))
    ras = ref.Rasterer(K, (w, h))
    out = {"K": _np(K), "width": w, "height": h, "coords": _np(sp), "normals": _np(sn), "pose": _np(pose), "bg": _np(bg)}
    for prim in ("circle", "circle_opt", "disc"):
        for use_bg in (False, True):
            coords = sp.clone().requires_grad_(True)
            r = ras(coords, sn, sn, pose, rot="dcm", primitives=prim, bg=(bg if use_bg else None), output_mask=True,
                    output_nocs=True, output_points=False)
            (g,) = torch.autograd.grad((r["color"] * bg).sum() + r["mask"].sum(), coords)
            tag = f"{prim}_{'bg' if use_bg else 'nobg'}"
            out[tag + "_color"], out[tag + "_mask"], out[tag + "_g_coords"] = _np(r["color"]), _np(r["mask"]), _np(g)
    np.savez_compressed(os.path.join(GOLDEN, "raster_primitives.npz"), **out)


def golden_losses(ref):
    prior = P.load_prior(STOCK_PRIOR)
    sc = scenes.make_scene(prior, size=32, density=20, n_lidar=150, seed=5)
    opt = ref.Optimizer({k: v.copy() for k, v in sc["init"].items()}, torch.device("cpu"), sc["weights"])
    opt.device, opt.precision = torch.device("cpu"), torch.float32
    # 3D: query = oracle-rendered front points of the initial state
    pts = O.lattice(20)
    st = O.RefineState.create(**sc["init"])
    out = O.iteration_losses(prior, pts, torch.tensor(sc["K"]), 32, 32, st, torch.tensor(sc["nocs_pred"]), sc["lidar"],
                             0.3, 0.5)
    xyzf = out["render"]["xyzf"].detach().clone().requires_grad_(True)
    lidar_s = (torch.Tensor(sc["lidar"]) / opt.params['scale']).detach().clone().requires_grad_(True)
    l3, _, _ = opt.compute_loss_3d(xyzf, lidar_s)
    g_xyzf, g_lidar = torch.autograd.grad(l3, [xyzf, lidar_s])
    color = out["render"]["color"].detach().clone().requires_grad_(True)
    target = out["target"].detach()
    l2 = opt.compute_loss_2d(color, target)
    (g_color,) = torch.autograd.grad(l2, [color])
    np.savez_compressed(
        os.path.join(GOLDEN, "losses.npz"), xyzf=_np(xyzf), lidar_scaled=_np(lidar_s), scale=_np(opt.params['scale']),
        loss3d=_np(l3), g_xyzf=_np(g_xyzf), g_lidar=_np(g_lidar), color=_np(color), target=_np(target),
        loss2d=_np(l2), g_color=_np(g_color))


def golden_refine(ref):
    """Trajectory of the reference's own Optimizer.optimize (5 iterations, 32x32, D=20)."""
    dec, L = ref.setup_dsdf(STOCK_PRIOR, precision=torch.float32)
    prior = P.load_prior(STOCK_PRIOR)
    sc = scenes.make_scene(prior, size=32, density=20, n_lidar=150, seed=5)
    params = {k: v.copy() for k, v in sc["init"].items()}
    opt = ref.Optimizer(params, torch.device("cpu"), sc["weights"])
    ref.grid_module.grads.clear()
    grid = ref.Grid3D(20)
    traj = []
    for it in range(5):
        opt.optimize(1, torch.tensor(sc["nocs_pred"]), sc["lidar"], dec, grid, torch.tensor(sc["K"]),
                     sc["crop_size"], viz_type=None)
        traj.append(np.concatenate([_np(params[k]).reshape(-1) for k in ("yaw", "trans", "scale", "latent")]))
    out = {"traj": np.stack(traj), "K": sc["K"], "crop_size": np.asarray(sc["crop_size"]), "density": 20,
           "nocs_pred": sc["nocs_pred"], "lidar": sc["lidar"], "w2d": 0.3, "w3d": 0.5}
    for k, v in sc["init"].items():
        out["init_" + k] = v
    np.savez_compressed(os.path.join(GOLDEN, "refine_traj.npz"), **out)


def main():
    if not ref_harness.available():
        sys.exit("reference tree not found - golden files can only be generated in the build container")
    if not os.path.isfile(STOCK_PRIOR):
        sys.exit("train the synthetic prior first: python -m oracle.prior")
    torch.manual_seed(1)
    np.random.seed(1)
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_harness.load()
    golden_lattice(ref)
    golden_decoders(ref)
    golden_stock(ref)
    golden_raster(ref)
    golden_primitives(ref)
    golden_losses(ref)
    golden_refine(ref)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
