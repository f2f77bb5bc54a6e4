"""Pins the oracle restatement against outputs of the UNMODIFIED reference.

The golden files were written by ``python -m oracle.make_golden`` in the build
container (reference imported from /root/reference).  Tolerances: the restatement
uses the same fp32 torch ops in a different order, so agreement is at the fp32
rounding floor (SURVEY.md Appendix B measured 1e-7 abs for sdf, 1e-6 for maps).
"""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import prior as P
from oracle import sdf_oracle as O


def _load(golden_dir, name):
    path = os.path.join(golden_dir, name)
    if not os.path.isfile(path):
        pytest.skip(f"{name} not generated")
    return np.load(path)


def test_lattice_bit_exact(golden_dir):
    g = _load(golden_dir, "lattice.npz")
    for d in (8, 9, 12):
        assert np.array_equal(O.lattice(d).numpy(), g[f"points_{d}"]), d
    for d in (30, 40, 41):
        h = hashlib.sha256(O.lattice(d).numpy().tobytes()).digest()
        assert np.array_equal(np.frombuffer(h, dtype=np.uint8), g[f"sha256_{d}"]), d


@pytest.mark.parametrize("name", ["wn_skip", "layernorm", "xyz_in_all", "use_tanh", "latent8", "latent256"])
def test_decoder_variants(golden_dir, name):
    g = _load(golden_dir, f"decoder_{name}.npz")
    spec = O.DecoderSpec.from_json(json.loads(bytes(g["spec_json"]).decode()))
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    params = O.params_from_state_dict(spec, sd)
    inp = torch.from_numpy(g["inputs"]).requires_grad_(True)
    sdf = O.decoder_forward(params, inp)
    (grad,) = torch.autograd.grad(sdf.sum(), inp)
    assert np.abs(sdf.detach().numpy() - g["sdf"]).max() < 2e-6
    assert np.abs(grad.numpy() - g["dinput"]).max() < 2e-5 * max(1.0, np.abs(g["dinput"]).max())


def test_stock_surface(golden_dir, stock_prior_path):
    g = _load(golden_dir, "stock_surface_d16.npz")
    prior = P.load_prior(stock_prior_path)
    pts = O.lattice(16)
    sdf, nrm, _ = O.sdf_and_normals(prior, torch.from_numpy(g["latent_unit"]), pts)
    assert np.abs(sdf.detach().numpy() - g["sdf"]).max() < 1e-6
    sp, nocs, sn, keep = O.surface_points(pts, sdf.detach(), nrm)
    assert np.array_equal(keep.numpy(), g["keep"])
    assert np.abs(sp.numpy() - g["surf_pts"]).max() < 1e-6
    assert np.abs(nocs.numpy() - g["surf_nocs"]).max() < 1e-6
    assert np.abs(sn.numpy() - g["surf_nrm"]).max() < 2e-5


@pytest.mark.parametrize("tag,rot", [("dcm_45x22", "dcm"), ("dcm_32x32", "dcm"), ("quat_40x30", "quat"),
                                     ("dcm_64x64", "dcm"), ("dcm_256x256", "dcm")])
def test_raster_maps_and_gradients(golden_dir, tag, rot):
    from tests import helpers as H
    g = _load(golden_dir, f"raster_{tag}.npz")
    w, h = int(g["width"]), int(g["height"])
    large = "cot_color" not in g.files          # 64x64 / 256x256 (reference tiled): recomputable cotangents
    coords = torch.from_numpy(g["coords"]).requires_grad_(True)
    normals = torch.from_numpy(g["normals"]).requires_grad_(True)
    pose = torch.from_numpy(g["pose"]).requires_grad_(True)
    r = O.render(torch.from_numpy(g["K"]), w, h, coords, normals, normals, pose, rot=rot, output_nocs=True,
                 tile_rows=32 if large else None)
    scalar = 0
    for i, k in enumerate(("color", "mask", "depth", "normals")):
        ref = g["r_" + k]
        assert np.abs(r[k].detach().numpy() - ref).max() < 2e-5 * max(1.0, np.abs(ref).max()), k
        cot = H.pattern_cotangent(tuple(ref.shape), i) if large else torch.from_numpy(g["cot_" + k])
        scalar = scalar + (r[k] * cot).sum()
    if rot == "dcm":
        assert np.abs(r["xyzf"].detach().numpy() - g["p_xyzf"]).max() < 1e-6
        assert np.abs(r["rgbf"].detach().numpy() - g["p_rgbf"]).max() < 1e-6
        assert np.abs(r["xyz"].detach().numpy() - g["p_xyz"]).max() < 1e-6
        cot = H.pattern_cotangent(tuple(g["p_xyzf"].shape), 7) if large else torch.from_numpy(g["cot_xyzf"])
        scalar = scalar + (r["xyzf"] * cot).sum()
    gc, gn, gp = torch.autograd.grad(scalar, [coords, normals, pose])
    for ours, ref in ((gc, g["g_coords"]), (gn, g["g_normals"]), (gp, g["g_pose"])):
        assert np.abs(ours.numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())


def test_losses(golden_dir):
    g = _load(golden_dir, "losses.npz")
    xyzf = torch.from_numpy(g["xyzf"]).requires_grad_(True)
    lidar = torch.from_numpy(g["lidar_scaled"]).requires_grad_(True)
    l3 = O.loss_3d(xyzf, lidar, float(g["scale"][0]))
    gx, gl = torch.autograd.grad(l3, [xyzf, lidar])
    assert abs(float(l3) - float(g["loss3d"])) < 1e-6
    assert np.abs(gx.numpy() - g["g_xyzf"]).max() < 1e-6
    assert np.abs(gl.numpy() - g["g_lidar"]).max() < 1e-6
    color = torch.from_numpy(g["color"]).requires_grad_(True)
    target = torch.from_numpy(g["target"])
    for dense in (False, True):
        l2 = O.loss_2d(color, target, dense=dense)
        (gcol,) = torch.autograd.grad(l2, [color])
        assert abs(float(l2) - float(g["loss2d"])) < 1e-6, dense
        assert np.abs(gcol.numpy() - g["g_color"]).max() < 1e-6, dense


def test_refine_trajectory(golden_dir, stock_prior_path):
    """Five iterations of the reference's Optimizer.optimize vs the oracle loop."""
    g = _load(golden_dir, "refine_traj.npz")
    prior = P.load_prior(stock_prior_path)
    pts = O.lattice(int(g["density"]))
    st = O.RefineState.create(g["init_yaw"], g["init_trans"], g["init_scale"], g["init_latent"])
    h, w = [int(v) for v in g["crop_size"]]
    traj = []
    for _ in range(g["traj"].shape[0]):
        O.refine_iteration(prior, pts, torch.from_numpy(g["K"]), w, h, st, torch.from_numpy(g["nocs_pred"]),
                           g["lidar"], float(g["w2d"]), float(g["w3d"]))
        p = st.as_numpy()
        traj.append(np.concatenate([p[k].reshape(-1) for k in ("yaw", "trans", "scale", "latent")]))
    traj = np.stack(traj)
    assert np.abs(traj - g["traj"]).max() < 2e-5, np.abs(traj - g["traj"]).max(0)


def _oracle_steps(prior, sc, steps):
    pts = O.lattice(sc["density"])
    st = O.RefineState.create(**sc["init"])
    h, w = sc["crop_size"]
    traj, losses = [], []
    for _ in range(steps):
        out = O.refine_iteration(prior, pts, torch.from_numpy(sc["K"]), w, h, st, torch.from_numpy(sc["nocs_pred"]),
                                 sc["lidar"], sc["weights"]["2d"], sc["weights"]["3d"])
        losses.append(float(out["loss"]) if not out["skip"] else float("nan"))
        p = st.as_numpy()
        traj.append(np.concatenate([p[k].reshape(-1) for k in ("yaw", "trans", "scale", "latent")]))
    return np.stack(traj), np.asarray(losses)


def test_refine_trajectory50_head(golden_dir, stock_prior_path):
    """The first steps of the reference's 50-step trajectory at cfg1 (64x64, D=40) vs the oracle loop, with the
    losses the reference printed (the GPU test follows all 50 steps)."""
    from tests import helpers as H
    g = _load(golden_dir, "refine_traj50_64.npz")
    sc = H.scene_from_golden(g)
    traj, losses = _oracle_steps(P.load_prior(stock_prior_path), sc, 4)
    assert np.abs(traj - g["traj"][:4]).max() < 2e-5, np.abs(traj - g["traj"][:4]).max(0)
    assert np.allclose(losses, g["losses"][:4, 2], rtol=1e-4)
    # sanity of the file itself: 50 steps, monotone-ish descent, nothing skipped
    assert g["traj"].shape == (50, 8) and np.isfinite(g["losses"]).all() and g["losses"][-1, 2] < 0.7 * g["losses"][0, 2]


def test_refine_ragged(golden_dir, stock_prior_path):
    """Three ragged detections (non-square crops, different LIDAR counts): the reference's Optimizer vs the oracle."""
    from tests import helpers as H
    g = _load(golden_dir, "refine_ragged.npz")
    prior = P.load_prior(stock_prior_path)
    for i in range(int(g["count"])):
        sc = H.scene_from_golden(g, f"d{i}_")
        assert sc["crop_size"][0] != sc["crop_size"][1]
        traj, losses = _oracle_steps(prior, sc, 3)
        assert np.abs(traj - g[f"d{i}_traj"][:3]).max() < 2e-5, i
        assert np.allclose(losses, g[f"d{i}_losses"][:3, 2], rtol=1e-4, equal_nan=True), i


def test_kitti_label_host_math(golden_dir):
    """roty_in_bev / alpha_in_bev / the label assembly of the product's host code against the reference's
    get_kitti_label outputs (the extents come from the golden file here; on the GPU they come from the kernels)."""
    from sdflabel_b200.utils.refinement import kitti_label_from_extents, roty_in_bev, alpha_in_bev
    g = _load(golden_dir, "kitti_label.npz")
    for i in range(int(g["count"])):
        f = lambda k: g[f"c{i}_{k}"]
        scale = f("scale").astype(np.float32)
        # un-scaled extents: the golden stores min / max of (points * scale); dividing is exact enough for 1e-6
        label, cam_T = kitti_label_from_extents(f("points_min") / scale, f("points_max") / scale, f("latent"), scale,
                                                f("trans"), f("yaw"), f("p_WC"), list(f("bbox")))
        assert np.abs(cam_T - f("cam_T")).max() < 1e-6
        global_T = np.linalg.inv(f("p_WC")) @ f("cam_T")
        assert abs(roty_in_bev(global_T) - float(f("rotation_y"))) < 1e-9
        assert abs(alpha_in_bev(global_T, float(f("rotation_y"))) - float(f("alpha"))) < 1e-9
        assert np.abs(np.asarray(label["dimensions"]) - f("dimensions")).max() < 1e-5
        assert np.abs(np.asarray(label["location"]) - f("location")).max() < 1e-5
        assert abs(label["alpha"] - float(f("alpha"))) < 1e-6


@pytest.mark.parametrize("prim", ["circle", "circle_opt", "disc"])
@pytest.mark.parametrize("use_bg", [False, True])
def test_other_primitives_and_background(golden_dir, prim, use_bg):
    """SURVEY 8(f) row 3 (oracle only in round 1): circle / circle_opt primitives and bg compositing against the
    reference Rasterer's colour and mask maps and point gradients."""
    g = np.load(os.path.join(golden_dir, "raster_primitives.npz"))
    w, h = int(g["width"]), int(g["height"])
    K, pose, bg = torch.from_numpy(g["K"]), torch.from_numpy(g["pose"]), torch.from_numpy(g["bg"])
    coords = torch.from_numpy(g["coords"]).requires_grad_(True)
    normals = torch.from_numpy(g["normals"])
    v, m, c, _ = O.to_camera(coords, normals, normals, pose, "dcm", True)
    if prim == "circle":
        wgt = O.circle_weights(K, w, h, v, add_bg=use_bg)
    elif prim == "circle_opt":
        wgt = O.circle_opt_weights(K, v, add_bg=use_bg)
    else:
        rays = O.pixel_rays(K, w, h)
        wgt = O.disc_weights_bg(rays, v, m) if use_bg else O.disc_weights(rays, v, m)
    if use_bg:
        color, mask = O.compose_bg(wgt, c, bg)
    else:
        color, mask, _, _ = O.compose(wgt, v, m, c, True)
    tag = f"{prim}_{'bg' if use_bg else 'nobg'}"
    assert np.abs(color.view(3, h, w).detach().numpy() - g[tag + "_color"]).max() < 1e-6
    assert np.abs(mask.view(1, h, w).detach().numpy() - g[tag + "_mask"]).max() < 1e-6
    (gc,) = torch.autograd.grad((color.view(3, h, w) * bg).sum() + mask.sum(), coords)
    ref = g[tag + "_g_coords"]
    assert np.abs(gc.numpy() - ref).max() < 5e-5 * max(1.0, np.abs(ref).max())   # fp32 accumulation order of the M x P sums
