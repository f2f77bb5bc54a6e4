"""Synthetic DeepSDF priors in the reference's checkpoint format (TEST INFRASTRUCTURE).

The KITTI-car prior (data/nets/deepsdf.{pt,json}) is not shipped with the
reference (README.md:27-28), so tests and benchmarks use synthetic priors:

* ``random_prior``  - seeded random weights for any spec (MLP-only parity tests).
* ``train_car_prior`` - the stock 8x512 / latent 3 / latent_in [4] / weight_norm
  spec fitted to a latent-parameterised analytic "car blob" so that the zero
  level set is car-sized inside [-1,1]^3 (random weights give a near-constant
  field with no surface).  The trained file is committed under ``assets/``.

Checkpoint layout (deepsdf/workspace.py:167-188): ``<path>.pt`` holds
``{"epoch", "model_state_dict"}`` with DataParallel ``module.`` key prefixes,
``<path>.json`` holds ``{"NetworkArch", "CodeLength", "NetworkSpecs"}``.
"""
from __future__ import annotations

import json
import os
import time
from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from .sdf_oracle import DecoderSpec, DecoderParams, params_from_state_dict, decoder_forward

STOCK_SPEC = DecoderSpec(
    latent_size=3,
    dims=[512] * 8,
    latent_in=(4,),
    norm_layers=tuple(range(8)),
    weight_norm=True,
    xyz_in_all=False,
    use_tanh=False,
    dropout=tuple(range(8)),
    dropout_prob=0.2,
    latent_dropout=False,
)


def _state_dict_skeleton(spec: DecoderSpec, gen: torch.Generator) -> Dict[str, torch.Tensor]:
    """nn.Linear-style init for every tensor the reference Decoder owns
    (deep_sdf_decoder_scale.py:44-75)."""
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, fan_in, fan_out, wn):
        bound = 1.0 / np.sqrt(fan_in)
        w = (torch.rand(fan_out, fan_in, generator=gen) * 2 - 1) * bound
        b = (torch.rand(fan_out, generator=gen) * 2 - 1) * bound
        if wn:
            sd[prefix + ".weight_g"] = w.norm(dim=1, keepdim=True)
            sd[prefix + ".weight_v"] = w
        else:
            sd[prefix + ".weight"] = w
        sd[prefix + ".bias"] = b

    for l, (i, o) in enumerate(spec.layer_dims()):
        linear(f"module.lin{l}", i, o, spec.uses_weight_norm(l))
        if spec.uses_layernorm(l):
            sd[f"module.bn{l}.weight"] = 1 + 0.1 * torch.randn(o, generator=gen)
            sd[f"module.bn{l}.bias"] = 0.1 * torch.randn(o, generator=gen)
    L = spec.latent_size
    linear("module.scale_net.0", L, 3, False)
    linear("module.scale_net.2", 3, 3, False)
    linear("module.scale_net.4", 3, 1, False)
    return sd


def random_prior(spec: DecoderSpec, seed: int = 1, gain: float = 1.6) -> Dict[str, torch.Tensor]:
    """Seeded random checkpoint.  ``gain`` widens the hidden layers a little so
    the field is not numerically constant."""
    gen = torch.Generator().manual_seed(seed)
    sd = _state_dict_skeleton(spec, gen)
    for k in list(sd):
        if k.startswith("module.lin") and (k.endswith("weight_g") or k.endswith(".weight")):
            sd[k] = sd[k] * gain
    return sd


def save_prior(path_pt: str, spec: DecoderSpec, state_dict: Dict[str, torch.Tensor], epoch: int = 0) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path_pt)), exist_ok=True)
    torch.save({"epoch": epoch, "model_state_dict": state_dict}, path_pt)
    with open(os.path.splitext(path_pt)[0] + ".json", "w") as f:
        json.dump(spec.to_json(), f, indent=1)


def load_prior(path_pt: str, dtype=torch.float32) -> DecoderParams:
    with open(os.path.splitext(path_pt)[0] + ".json") as f:
        spec = DecoderSpec.from_json(json.load(f))
    sd = torch.load(path_pt, map_location="cpu")["model_state_dict"]
    return params_from_state_dict(spec, sd).to(dtype)


# ----------------------------------------------------------------------------
# Analytic "car blob" family
# ----------------------------------------------------------------------------
def _ellipsoid_bound(x, centre, radii):
    q = x - centre
    k0 = (q / radii).norm(dim=-1)
    k1 = (q / (radii * radii)).norm(dim=-1).clamp_min(1e-9)
    return k0 * (k0 - 1.0) / k1


def car_blob_sdf(latent_unit: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """Approximate SDF of body + cabin ellipsoids whose radii depend on the
    unit latent.  x: (..., 3) in the object frame (x width, y height, z length)."""
    l0, l1, l2 = latent_unit[..., 0], latent_unit[..., 1], latent_unit[..., 2]
    body_r = torch.stack([0.36 + 0.06 * l0, 0.21 + 0.04 * l1, 0.78 + 0.10 * l2], dim=-1)
    body_c = torch.stack([torch.zeros_like(l0), -0.08 + 0.0 * l0, torch.zeros_like(l0)], dim=-1)
    cab_r = torch.stack([0.30 + 0.05 * l0, 0.20 + 0.06 * l1, 0.40 + 0.08 * l2], dim=-1)
    cab_c = torch.stack([torch.zeros_like(l0), 0.12 + 0.04 * l1, -0.08 + 0.1 * l0], dim=-1)
    return torch.minimum(_ellipsoid_bound(x, body_c, body_r), _ellipsoid_bound(x, cab_c, cab_r))


def train_car_prior(path_pt: str, steps: int = 600, seed: int = 1, batch_latents: int = 16,
                    pts_per_latent: int = 1024, lr: float = 5e-4, log_every: int = 50) -> None:
    """Fits the stock spec to ``car_blob_sdf`` with the DeepSDF clamped-L1 loss."""
    torch.manual_seed(seed)
    spec = STOCK_SPEC
    gen = torch.Generator().manual_seed(seed)
    sd = {k: v.clone().requires_grad_(True) for k, v in _state_dict_skeleton(spec, gen).items()}
    train_keys = [k for k in sd if k.startswith("module.lin")]
    opt = torch.optim.Adam([sd[k] for k in train_keys], lr=lr)
    t0 = time.time()
    for step in range(steps):
        lat = torch.nn.functional.normalize(torch.randn(batch_latents, 1, 3, generator=gen), dim=-1)
        uni = torch.rand(batch_latents, pts_per_latent // 2, 3, generator=gen) * 2.1 - 1.025
        # near-surface samples: points on the body/cabin ellipsoids plus noise
        u = torch.nn.functional.normalize(torch.randn(batch_latents, pts_per_latent // 2, 3, generator=gen), dim=-1)
        pick = torch.rand(batch_latents, pts_per_latent // 2, 1, generator=gen) < 0.6
        l0, l1, l2 = lat[..., 0:1], lat[..., 1:2], lat[..., 2:3]
        body = u * torch.cat([0.36 + 0.06 * l0, 0.21 + 0.04 * l1, 0.78 + 0.10 * l2], -1) + torch.tensor([0.0, -0.08, 0.0])
        cab = u * torch.cat([0.30 + 0.05 * l0, 0.20 + 0.06 * l1, 0.40 + 0.08 * l2], -1) + torch.cat(
            [torch.zeros_like(l0), 0.12 + 0.04 * l1, -0.08 + 0.1 * l0], -1)
        near = torch.where(pick, body, cab) + 0.03 * torch.randn(batch_latents, pts_per_latent // 2, 3, generator=gen)
        x = torch.cat([uni, near], dim=1)
        tgt = car_blob_sdf(lat.expand(-1, x.shape[1], -1), x).reshape(-1, 1)
        inp = torch.cat([lat.expand(-1, x.shape[1], -1), x], dim=-1).reshape(-1, 6)
        params = params_from_state_dict(spec, sd)
        pred = decoder_forward(params, inp)
        loss = (pred.clamp(-0.1, 0.1) - tgt.clamp(-0.1, 0.1)).abs().mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % log_every == 0 or step == steps - 1:
            print(f"[prior] step {step:4d} loss {loss.item():.5f}  ({time.time() - t0:.0f}s)", flush=True)
    save_prior(path_pt, spec, {k: v.detach().clone() for k, v in sd.items()}, epoch=steps)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "assets", "deepsdf_synth.pt"))
    ap.add_argument("--steps", type=int, default=600)
    a = ap.parse_args()
    train_car_prior(a.out, steps=a.steps)
