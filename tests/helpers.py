"""Shared helpers of the parity tests (the oracle is the checker, never the thing under test)."""
import json
import os

import numpy as np
import torch

from oracle import prior as P
from oracle import sdf_oracle as O


def our_decoder_from_state(spec: O.DecoderSpec, state_dict, device="cuda"):
    """Instantiates the product Decoder exactly the way setup_dsdf does."""
    from sdflabel_b200.deepsdf.networks.deep_sdf_decoder_scale import Decoder
    ns = spec.to_json()["NetworkSpecs"]
    dec = Decoder(spec.latent_size, **ns)
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    dec.load_state_dict(sd)
    return dec.to(device).eval()


def golden_decoder(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"decoder_{name}.npz"))
    spec = O.DecoderSpec.from_json(json.loads(bytes(g["spec_json"]).decode()))
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    return g, spec, sd


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())


def frac_within(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    return float(ok.mean())


def pattern_cotangent(shape, salt):
    """Same deterministic cotangent as oracle/make_golden_r2.py::pattern_cotangent (recomputed, not stored)."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    h = (idx * np.uint64(2654435761) + np.uint64(salt * 97 + 13)) % np.uint64(1 << 20)
    return torch.from_numpy((h.astype(np.float64) / float(1 << 20)).astype(np.float32).reshape(shape))


def scene_from_golden(g, prefix=""):
    """Scene dict (the arguments of Optimizer.optimize) from the fields oracle/make_golden*.py store."""
    f = lambda k: g[prefix + k]
    return {"K": f("K"), "crop_size": [int(v) for v in f("crop_size")], "density": int(f("density")),
            "nocs_pred": f("nocs_pred"), "lidar": f("lidar"),
            "weights": {"2d": float(f("w2d")), "3d": float(f("w3d"))},
            "init": {k: f("init_" + k) for k in ("yaw", "trans", "scale", "latent")}}


def camera_space(coords, normals, pose):
    """v = R p + t, m = R n in float64 (projection.py:49,58) for the threshold attribution below."""
    pose = np.asarray(pose, dtype=np.float64)
    R, t = pose[:3, :3], pose[:3, 3]
    return np.asarray(coords, np.float64) @ R.T + t, np.asarray(normals, np.float64) @ R.T


def threshold_margin(K, width, v, m, pixels, chunk=256):
    """For each flat pixel index j = y*W + x: how close the nearest surfel is to one of the splat's HARD
    decisions at that pixel - the disc edge ``diam - |v - g| > 0`` (primitives.py:215-226) or the ray cut-off
    ``abs(b) < 0.01`` (primitives.py:209-210).  A float32 implementation may legitimately decide such a surfel
    either way; SURVEY.md Appendix C (T5) asks that every out-of-tolerance pixel be attributed to one."""
    kinv = np.linalg.inv(np.asarray(K, dtype=np.float32)).astype(np.float64)
    pixels = np.asarray(pixels, dtype=np.int64).reshape(-1)
    out = np.empty(pixels.shape[0], dtype=np.float64)
    a = (m * v).sum(1)                                           # (M,)
    eps = float(np.finfo(np.float32).eps)
    for s in range(0, pixels.shape[0], chunk):
        px = pixels[s:s + chunk]
        pix = np.stack([px % width, px // width, np.ones_like(px)], 1).astype(np.float64)
        rays = pix @ kinv.T                                      # (P,3)
        b = m @ rays.T                                           # (M,P)
        cut = np.abs(np.abs(b) - 0.01)
        b = np.where(np.abs(b) < 0.01, eps, b)
        z = a[:, None] / b
        hit = rays[None, :, :] * z[:, :, None]
        gap = 0.04 - np.sqrt(((v[:, None, :] - hit) ** 2).sum(-1))
        # the cut-off only matters where the surfel is anywhere near the pixel's ray
        near = np.sqrt(((v[:, None, :] - rays[None] * v[:, None, 2:3]) ** 2).sum(-1)) < 0.2
        out[s:s + chunk] = np.minimum(np.abs(gap).min(0), np.where(near, cut, np.inf).min(0))
    return out


def ray_distance(K, width, v, pixels):
    """(P,) distance from each pixel's ray to the nearest of the points ``v`` (camera frame)."""
    kinv = np.linalg.inv(np.asarray(K, dtype=np.float32)).astype(np.float64)
    px = np.asarray(pixels, dtype=np.int64).reshape(-1)
    rays = np.stack([px % width, px // width, np.ones_like(px)], 1).astype(np.float64) @ kinv.T
    rays /= np.linalg.norm(rays, axis=1, keepdims=True)
    t = v @ rays.T                                               # (M,P) ray parameter of the closest approach
    d = np.linalg.norm(v[:, None, :] - t[:, :, None] * rays[None], axis=-1)
    return d.min(0)


def assert_maps_close(ours, ref, K, width, v, m, name, tol=1e-4, margin=2e-6, kink_points=None):
    """Every pixel within ``tol`` (relative to the map's maximum, north_star 1e-4) except pixels at which a surfel
    sits within ``margin`` of a hard threshold - or, when the two sides were given surfels of their OWN (the
    engine-level tests), pixels whose ray passes one of ``kink_points``: surfels whose normal legitimately differs
    between two fp32 evaluations of the decoder (a hidden unit on its ReLU kink), which tilts that
    surfel's disc.  Returns the number of attributed pixels."""
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    err = np.abs(ours - ref).reshape(ref.shape[0], -1).max(0)
    bad = np.nonzero(err > tol * max(1.0, np.abs(ref).max()))[0]
    if bad.size:
        mg = threshold_margin(K, width, v, m, bad)
        explained = mg < margin
        if kink_points is not None and len(kink_points):
            explained |= ray_distance(K, width, np.asarray(kink_points, np.float64), bad) < 0.06   # disc radius 0.04
        worst = float(mg[~explained].max()) if (~explained).any() else 0.0
        assert explained.all(), (f"{name}: {bad.size} pixels out of tolerance, {int((~explained).sum())} of them "
                                 f"unexplained (up to {worst:.2e} away from any hard threshold, max err {err.max():.2e})")
    return int(bad.size)


def min_abs_preactivation(p: "O.DecoderParams", inputs: torch.Tensor) -> np.ndarray:
    """Per input row: the smallest |pre-activation| over all hidden units (float64).  A ReLU whose input is
    this close to zero may switch between two fp32 evaluations of the same network and change d sdf / d x by a
    finite amount (every gradient outlier must be such a point)."""
    p = p.to(torch.float64)
    spec = p.spec
    inputs = inputs.to(torch.float64)
    xyz = inputs[:, -3:]
    x = inputs
    best = torch.full((inputs.shape[0],), float("inf"), dtype=torch.float64)
    n_lin = len(p.weight)
    for l in range(n_lin - 1):
        if l in spec.latent_in:
            x = torch.cat([x, inputs], dim=1)
        elif l != 0 and spec.xyz_in_all:
            x = torch.cat([x, xyz], dim=1)
        x = x @ p.weight[l].t() + p.bias[l]
        if spec.uses_layernorm(l):
            x = torch.nn.functional.layer_norm(x, (x.shape[1],), p.ln_weight[l], p.ln_bias[l], 1e-5)
        best = torch.minimum(best, x.abs().min(dim=1)[0])
        x = torch.relu(x)
    return best.numpy()


def assert_grad_rows_close(ours, ref, params, inputs, name, rtol=1e-4, frac=0.999, kink=2e-5):
    """T3: rows (points) whose gradient differs by more than rtol (relative to the largest gradient entry)
    are at most 1 - frac of all rows, and each of them has a hidden unit within ``kink`` of its ReLU kink."""
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max()
    row_err = np.abs(ours - ref).max(1) / scale
    bad = np.nonzero(row_err > rtol)[0]
    assert bad.size <= (1.0 - frac) * ref.shape[0] + 1, (name, bad.size, ref.shape[0], row_err.max())
    if bad.size:
        pre = min_abs_preactivation(params, inputs[bad])
        assert pre.max() < kink, (name, "gradient outlier away from any ReLU kink", pre.max(), row_err[bad].max())
    return int(bad.size)
