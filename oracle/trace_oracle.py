"""Torch restatement of trace mode (TEST INFRASTRUCTURE): sphere tracing of the DeepSDF
decoder per pixel ray, SURVEY.md Appendix A8.  There is no reference implementation of this
renderer (the reference only has the surfel splat), so this file *is* the specification the
CUDA trace kernels are checked against; it uses the same decoder restatement as the rest of
the oracle.  Gradients come from autograd through one Newton step at the converged hit
(tau = tau* - f(x*)/(grad f . d) with the denominator detached), which is exactly the
implicit-function derivative of f(l, o + tau d) = 0.
"""
from __future__ import annotations

import torch

from . import sdf_oracle as O

BOX_LO, BOX_HI = -1.0, 1.025


def rays(K, width, height, pose):
    """Object-frame origin o (3,), unit directions d (P,3) and unit camera rays rn (P,3)."""
    r = O.pixel_rays(K, width, height)
    rn = r / r.norm(dim=1, keepdim=True)
    R, t = pose[:3, :3], pose[:3, 3]
    o = -(R.t() @ t)
    d = rn @ R                       # rows: R^T rn
    return o, d, rn


def trace(params: O.DecoderParams, latent_unit, K, width, height, pose, max_steps=64, eps=1e-4):
    """Returns dict(depth (1,H,W), normals (3,H,W), nocs (3,H,W), mask (1,H,W), slope (1,H,W)); depth and nocs
    are differentiable with respect to pose and latent_unit.  The march stops anywhere inside |f| < eps, so the ray
    parameter of a hit is only defined to eps / slope: an implementation that stops elsewhere in that band (or polishes
    the root) is equally right, and the tests compare with that per-ray tolerance."""
    P = width * height
    with torch.no_grad():
        o, d, rn = rays(K, width, height, pose)
        inv = 1.0 / d
        ta, tb = (BOX_LO - o) * inv, (BOX_HI - o) * inv
        t0 = torch.minimum(ta, tb).max(dim=1)[0].clamp(min=0.0)
        t1 = torch.maximum(ta, tb).min(dim=1)[0]
        active = t0 <= t1
        tau = t0.clone()
        hit = torch.zeros(P, dtype=torch.bool)
        for step in range(max_steps):
            idx = active.nonzero().squeeze(1)
            if idx.numel() == 0:
                break
            x = o + tau[idx, None] * d[idx]
            f = O.decoder_forward(params, torch.cat([latent_unit.detach().expand(idx.numel(), -1), x], 1)).squeeze(1)
            conv = f.abs() < eps
            hit[idx[conv]] = True
            active[idx[conv]] = False
            go = ~conv
            tau[idx[go]] = tau[idx[go]] + f[go]
            out = go & ((tau[idx] > t1[idx]) | (tau[idx] < 0))
            active[idx[out]] = False
            if step == max_steps - 1:
                active[:] = False
    hid = hit.nonzero().squeeze(1)
    return render_hits(params, latent_unit, K, width, height, pose, tau, hid)


def render_hits(params: O.DecoderParams, latent_unit, K, width, height, pose, tau, hid):
    """The maps of `trace` for given hit pixels `hid` (flat indices) and ray parameters `tau` ((P,), constants):
    values at o + tau d, gradients by one differentiable Newton step (= the implicit-function derivative).  Also the
    way to evaluate the specification's gradients at the hit points ANOTHER march found."""
    P = width * height
    o, d, rn = rays(K, width, height, pose)                      # with the graph this time
    depth = torch.zeros(P, dtype=K.dtype)
    nocs = torch.zeros(3, P, dtype=K.dtype)
    nmap = torch.zeros(3, P, dtype=K.dtype)
    mask = torch.zeros(P, dtype=K.dtype)
    slope = torch.zeros(P, dtype=K.dtype)          # |grad f . d| at the hit: eps / slope is the width of the stopping band in tau
    if hid.numel():
        x0 = (o + tau[hid, None] * d[hid])
        xg = x0.detach().requires_grad_(True)
        f_g = O.decoder_forward(params, torch.cat([latent_unit.detach().expand(hid.numel(), -1), xg], 1))
        (G,) = torch.autograd.grad(f_g.sum(), xg)
        f = O.decoder_forward(params, torch.cat([latent_unit.expand(hid.numel(), -1), x0], 1)).squeeze(1)
        denom = (G * d[hid]).sum(1).detach()
        tau_star = tau[hid] - (f - f.detach()) / denom            # value tau*, gradient = implicit derivative
        xs = o + tau_star[:, None] * d[hid]
        depth = depth.index_put((hid,), tau_star * rn[hid, 2])
        sign = torch.tensor([-1.0, 1.0, 1.0], dtype=K.dtype)
        nocs = nocs.index_put((torch.arange(3)[:, None], hid[None, :]), ((xs * sign + 1) / 2).t())
        n_cam = (G / G.norm(dim=1, keepdim=True)) @ pose[:3, :3].detach().t()
        nmap = nmap.index_put((torch.arange(3)[:, None], hid[None, :]), ((n_cam + 1) / 2).t())
        mask = mask.index_put((hid,), torch.ones(hid.numel(), dtype=K.dtype))
        slope = slope.index_put((hid,), denom.abs())
    return {"depth": depth.view(1, height, width), "nocs": nocs.view(3, height, width),
            "normals": nmap.view(3, height, width), "mask": mask.view(1, height, width), "hits": hid,
            "slope": slope.view(1, height, width)}
