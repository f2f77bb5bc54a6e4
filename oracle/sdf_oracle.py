"""Torch-CPU restatement of the sdflabel render/refine hot path (TEST INFRASTRUCTURE).

See ``oracle/__init__.py`` for the rules.  All functions are dtype-parametric
(fp32 to mirror the reference, fp64 to measure its noise floor) and written from
the formulas in SURVEY.md Appendix A, not from the reference's code.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------
# Magic numbers of the reference (SURVEY.md section 5 "Config / flags")
# ----------------------------------------------------------------------------
BAND = 0.03            # grid.py:43
DISC_RADIUS = 0.04     # rasterer.py:103 ("diam", used as a radius in primitives.py:220)
DEPTH_GAIN = 150.0     # primitives.py:172
RAY_CUTOFF = 0.01      # primitives.py:210
NN_RADIUS = 0.2        # optimizer.py:166
WIN_RADIUS = 5.0       # optimizer.py:200 (diam)
NOCS_THR = 1.0         # optimizer.py:200 (threshold_nocs)


# ----------------------------------------------------------------------------
# Lattice  (sdfrenderer/grid.py:22-41)
# ----------------------------------------------------------------------------
def lattice(density: int, dtype=torch.float32) -> torch.Tensor:
    """(D^3, 3) sample lattice in [-1, 1]^3, z fastest; every odd *row* of the
    flattened array is shifted by half a cell in x and y (grid.py:38).  The
    reference builds it in float64 and rounds to float32 (grid.py:39); a higher
    ``dtype`` keeps those float32 values."""
    d = int(density)
    axis = -1.0 + np.arange(d, dtype=np.float64) * (2.0 / (d - 1)) if d > 1 else np.array([-1.0])
    # np.mgrid[-1:1:d*1j] == start + arange(d) * step with step = (stop-start)/(d-1)
    ix, iy, iz = np.meshgrid(np.arange(d), np.arange(d), np.arange(d), indexing="ij")
    pts = np.stack([axis[ix], axis[iy], axis[iz]], axis=-1).reshape(-1, 3)
    shift = (axis.max() - axis.min()) / d / 2.0
    pts[1::2, :2] += shift
    return torch.from_numpy(pts.astype(np.float32)).to(dtype)


# ----------------------------------------------------------------------------
# DeepSDF decoder  (deepsdf/networks/deep_sdf_decoder_scale.py:10-114)
# ----------------------------------------------------------------------------
@dataclass
class DecoderSpec:
    """The subset of ``NetworkSpecs`` that changes the function in eval mode."""
    latent_size: int
    dims: List[int]
    latent_in: Tuple[int, ...] = ()
    norm_layers: Tuple[int, ...] = ()
    weight_norm: bool = False
    xyz_in_all: bool = False
    use_tanh: bool = False
    dropout: Optional[Tuple[int, ...]] = None     # eval no-op
    dropout_prob: float = 0.0                     # eval no-op
    latent_dropout: bool = False                  # eval no-op

    @staticmethod
    def from_json(specs: dict) -> "DecoderSpec":
        ns = dict(specs["NetworkSpecs"])
        ns.pop("samples_per_scene", None)  # workspace.py:174
        return DecoderSpec(
            latent_size=int(specs["CodeLength"]),
            dims=list(ns["dims"]),
            latent_in=tuple(ns.get("latent_in", ())),
            norm_layers=tuple(ns.get("norm_layers", ())),
            weight_norm=bool(ns.get("weight_norm", False)),
            xyz_in_all=bool(ns.get("xyz_in_all", False) or False),
            use_tanh=bool(ns.get("use_tanh", False)),
            dropout=tuple(ns["dropout"]) if ns.get("dropout") is not None else None,
            dropout_prob=float(ns.get("dropout_prob", 0.0)),
            latent_dropout=bool(ns.get("latent_dropout", False)),
        )

    def to_json(self) -> dict:
        return {
            "NetworkArch": "deep_sdf_decoder_scale",
            "CodeLength": self.latent_size,
            "NetworkSpecs": {
                "dims": list(self.dims),
                "dropout": list(self.dropout) if self.dropout is not None else None,
                "dropout_prob": self.dropout_prob,
                "norm_layers": list(self.norm_layers),
                "latent_in": list(self.latent_in),
                "xyz_in_all": self.xyz_in_all,
                "use_tanh": self.use_tanh,
                "latent_dropout": self.latent_dropout,
                "weight_norm": self.weight_norm,
            },
        }

    # layer table -----------------------------------------------------------
    def layer_dims(self) -> List[Tuple[int, int]]:
        """(in, out) of every Linear, following deep_sdf_decoder_scale.py:29-52."""
        full = [self.latent_size + 3] + list(self.dims) + [1]
        n = len(full)
        out = []
        for l in range(n - 1):
            if (l + 1) in self.latent_in:
                o = full[l + 1] - full[0]
            else:
                o = full[l + 1]
                if self.xyz_in_all and l != n - 2:
                    o -= 3
            out.append((full[l], o))
        return out

    def uses_layernorm(self, l: int) -> bool:
        return (not self.weight_norm) and (self.norm_layers is not None) and (l in self.norm_layers)

    def uses_weight_norm(self, l: int) -> bool:
        return self.weight_norm and (l in self.norm_layers)


def fold_weight_norm(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """W[r,:] = g[r] * v[r,:] / ||v[r,:]||  (torch weight_norm, dim=0)."""
    return v * (g / v.norm(dim=1, keepdim=True))


@dataclass
class DecoderParams:
    """Effective (folded) parameters, one entry per Linear."""
    spec: DecoderSpec
    weight: List[torch.Tensor]
    bias: List[torch.Tensor]
    ln_weight: Dict[int, torch.Tensor] = field(default_factory=dict)
    ln_bias: Dict[int, torch.Tensor] = field(default_factory=dict)

    def to(self, dtype) -> "DecoderParams":
        return DecoderParams(
            self.spec,
            [w.to(dtype) for w in self.weight],
            [b.to(dtype) for b in self.bias],
            {k: v.to(dtype) for k, v in self.ln_weight.items()},
            {k: v.to(dtype) for k, v in self.ln_bias.items()},
        )


def params_from_state_dict(spec: DecoderSpec, sd: Dict[str, torch.Tensor]) -> DecoderParams:
    """Reads the reference checkpoint layout (workspace.py:176-180: keys carry a
    ``module.`` prefix; weight-normed layers store ``weight_g``/``weight_v``)."""
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    ws, bs, lw, lb = [], [], {}, {}
    for l, _ in enumerate(spec.layer_dims()):
        if f"lin{l}.weight_v" in sd:
            ws.append(fold_weight_norm(sd[f"lin{l}.weight_v"], sd[f"lin{l}.weight_g"]))
        else:
            ws.append(sd[f"lin{l}.weight"])
        bs.append(sd[f"lin{l}.bias"])
        if f"bn{l}.weight" in sd:
            lw[l], lb[l] = sd[f"bn{l}.weight"], sd[f"bn{l}.bias"]
    return DecoderParams(spec, ws, bs, lw, lb)


def decoder_forward(p: DecoderParams, inputs: torch.Tensor) -> torch.Tensor:
    """(N, L+3) -> (N, 1).  deep_sdf_decoder_scale.py:78-107, eval mode."""
    spec = p.spec
    n_lin = len(p.weight)
    xyz = inputs[:, -3:]
    x = inputs
    for l in range(n_lin):
        if l in spec.latent_in:
            x = torch.cat([x, inputs], dim=1)
        elif l != 0 and spec.xyz_in_all:
            x = torch.cat([x, xyz], dim=1)
        x = x @ p.weight[l].t() + p.bias[l]
        if l == n_lin - 1 and spec.use_tanh:
            x = torch.tanh(x)
        if l < n_lin - 1:
            if spec.uses_layernorm(l):
                x = torch.nn.functional.layer_norm(x, (x.shape[1],), p.ln_weight[l], p.ln_bias[l], 1e-5)
            x = torch.relu(x)
    return torch.tanh(x)


# ----------------------------------------------------------------------------
# Surface extraction  (sdfrenderer/grid.py:43-71)
# ----------------------------------------------------------------------------
def sdf_and_normals(p: DecoderParams, latent_unit: torch.Tensor, pts: torch.Tensor):
    """Returns (sdf (N,1) still attached to ``latent_unit``, unit normals (N,3)
    as constants, raw gradient (N,3)).  The reference gets the gradient through
    a tensor hook on the lattice (grid.py:55-56) and normalises it in place with
    a detached norm (grid.py:57-58); the result is a constant in the graph."""
    x = pts.detach().clone().requires_grad_(True)
    inp = torch.cat([latent_unit.expand(x.shape[0], -1), x], dim=1)
    sdf = decoder_forward(p, inp)
    # When the decoder tensors require grad (bench.py's reference-cost mode) the backward also
    # produces dW for every layer, as the reference's sdf.sum().backward() does (grid.py:55).
    extra = [t for t in list(p.weight) + list(p.bias) if t.requires_grad]
    g = torch.autograd.grad(sdf.sum(), [x] + extra, retain_graph=True)[0]
    nrm = g / g.norm(dim=1, keepdim=True)
    return sdf, nrm.detach(), g.detach()


def surface_points(pts: torch.Tensor, sdf: torch.Tensor, nrm: torch.Tensor, thr: float = BAND):
    """grid.py:61-67: p = g - f*n; keep |f| < thr, ascending index order."""
    proj = pts - sdf * nrm
    keep = (sdf.abs() < thr).squeeze(1)
    return proj[keep], (proj[keep] + 1) / 2, nrm[keep], keep


# ----------------------------------------------------------------------------
# Pose helpers  (utils/refinement.py:108-125, optimizer.py:87-90, utils_rasterer.py:6-24)
# ----------------------------------------------------------------------------
def yaw_pose(yaw: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    """4x4 render pose: diag(1,-1,1) * R_y(yaw), translation overwritten after
    the row flip (optimizer.py:87-90)."""
    c, s = torch.cos(yaw).reshape(()), torch.sin(yaw).reshape(())
    z, o = torch.zeros_like(c), torch.ones_like(c)
    rot = torch.stack([torch.stack([c, z, s]), torch.stack([z, -o, z]), torch.stack([-s, z, c])])
    top = torch.cat([rot, trans.reshape(3, 1)], dim=1)
    return torch.cat([top, torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=top.dtype)], dim=0)


def quat_rotate(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """v + 2 (q_w (q_xyz x v) + q_xyz x (q_xyz x v)); q is not normalised."""
    qv = q[1:].expand_as(v)
    uv = torch.linalg.cross(qv, v, dim=1)
    uuv = torch.linalg.cross(qv, uv, dim=1)
    return v + 2 * (q[0] * uv + uuv)


# ----------------------------------------------------------------------------
# Projection  (renderer/projection.py:7-101, 104-199)
# ----------------------------------------------------------------------------
def to_camera(points, normals, colors, pose, rot: str, output_nocs: bool):
    """Returns camera-space points v, normals m, colours c and the front-facing
    mask.  DCM path negates x of the NOCS colours (projection.py:53-55) and
    filters by m.v < 0 (61-70); the quaternion path does neither by default
    (projection.py:105,149,160)."""
    if rot == "dcm":
        rt = pose[:3]
        m = normals @ rt[:, :3].t()
        v = points @ rt[:, :3].t() + rt[:, 3]
        if output_nocs:
            c = points * points.new_tensor([-1.0, 1.0, 1.0])
        else:
            c = colors
        front = (m * v).sum(1) < 0
    elif rot == "quat":
        q, t = pose[:4], pose[4:]
        m = quat_rotate(q, normals)
        v = quat_rotate(q, points) + t
        c = points if output_nocs else colors
        front = None
    else:
        raise ValueError(rot)
    return v, m, c, front


def project_pixels(K, v, res_xy):
    """projection.py:88-93 (only consumed by the circle primitives)."""
    eps = torch.finfo(K.dtype).eps
    h = v @ K.t()
    xy = h[:, :2] / (h[:, 2:] + eps)
    return torch.stack([xy[:, 0].clamp(-1, res_xy[0]), xy[:, 1].clamp(-1, res_xy[1])], dim=1)


# ----------------------------------------------------------------------------
# Disc splat  (renderer/primitives.py:165-243 with diam=0.04, softclamp=False,
# add_bg=False; composition renderer/rasterer.py:113-144)
# ----------------------------------------------------------------------------
def pixel_rays(K: torch.Tensor, width: int, height: int, rows: Optional[Tuple[int, int]] = None):
    """r_j = K^-1 [x, y, 1]; K inverted in fp32 (primitives.py:204); pixel
    centres are integer coordinates, j = y*W + x (rasterer.py:25-27)."""
    y0, y1 = rows if rows is not None else (0, height)
    yy, xx = torch.meshgrid(torch.arange(y0, y1), torch.arange(width), indexing="ij")
    pix = torch.stack([xx.reshape(-1), yy.reshape(-1), torch.ones(xx.numel(), dtype=torch.long)], dim=1)
    kinv = K.float().inverse().to(K.dtype)
    return pix.to(K.dtype) @ kinv.t()


def disc_weights(rays, v, m, radius=DISC_RADIUS, gain=DEPTH_GAIN):
    """(M, P) depth-softmax weights w_ij of surfel i at pixel j."""
    dtype = v.dtype
    eps = torch.finfo(dtype).eps
    a = (m * v).sum(1, keepdim=True)                      # (M,1)   n.v
    b = m @ rays.t()                                      # (M,P)   n.r
    b = torch.where(b.abs() < RAY_CUTOFF, torch.full_like(b, eps).detach(), b)
    z = a / b                                             # ray / tangent-plane hit depth
    hit = rays.unsqueeze(0) * z.unsqueeze(-1)             # (M,P,3)
    gap = radius - (v.unsqueeze(1) - hit).pow(2).sum(-1).sqrt()
    inside = (gap.clamp(min=0) > 0).detach()
    zeta = -z * inside.to(dtype)
    nu = zeta.norm(dim=0).detach()
    score = (zeta / (nu.unsqueeze(0) + eps) + 1).clamp(min=0) * gain
    score = score.masked_fill(~inside, torch.finfo(dtype).min)
    return torch.softmax(score, dim=0) * inside.to(dtype)


def compose(w, v, m, c, output_nocs: bool):
    """rasterer.py:113-144.  Returns flat maps (3,P),(1,P),(1,P),(3,P)."""
    col = (c + 1) / 2 if output_nocs else c
    color = (w.unsqueeze(1) * col.unsqueeze(-1)).sum(0).clamp(max=1)
    mask = w.sum(0, keepdim=True).clamp(max=1)
    depth = (w * v[:, 2:3]).sum(0, keepdim=True)
    normals = (w.unsqueeze(1) * ((m + 1) / 2).unsqueeze(-1)).sum(0).clamp(max=1)
    return color, mask, depth, normals


def render(K, width, height, points, normals, colors, pose, rot="dcm", output_nocs=True,
           tile_rows: Optional[int] = None):
    """Full ``Rasterer.forward`` (disc primitive, no background).  ``tile_rows``
    evaluates the per-pixel math over row tiles (bit-identical)."""
    v, m, c, front = to_camera(points, normals, colors, pose, rot, output_nocs)
    step = tile_rows or height
    parts = []
    for y0 in range(0, height, step):
        rays = pixel_rays(K, width, height, (y0, min(height, y0 + step)))
        w = disc_weights(rays, v, m)
        parts.append(compose(w, v, m, c, output_nocs))
    color, mask, depth, nrm = [torch.cat([p[i] for p in parts], dim=1) for i in range(4)]
    out = {
        "color": color.view(3, height, width),
        "mask": mask.view(1, height, width),
        "depth": depth.view(1, height, width),
        "normals": nrm.view(3, height, width),
        "xyz": v,
        "rgb": (c + 1) / 2,
    }
    if front is not None:
        out["xyzf"] = v[front]
        out["rgbf"] = (c[front] + 1) / 2
        out["front"] = front
    return out


# ----------------------------------------------------------------------------
# The other two primitives and background compositing (only the
# demo CLI sdfrenderer/main.py reaches them, always with bg=None).  Oracle only in
# round 1: the product raises NotImplementedError for them.
# ----------------------------------------------------------------------------
def _depth_scores(v, gain):
    """primitives.py:52-56 / 143-146: per-point score from the camera depth alone."""
    eps = torch.finfo(v.dtype).eps
    z = -v[:, 2:]
    nu = z.norm(p=2, dim=0).detach()
    return (z / (nu.unsqueeze(0) + eps) + 1).clamp(min=0) * gain          # (M,1)


def circle_weights(K, width, height, v, add_bg=False, diam=0.02, gain=100.0, soft=3.0):
    """``inside_circle`` as ``Rasterer`` calls it (primitives.py:4-68, rasterer.py:93-96): the sigmoid "soft clamp" is
    only tested for > 0, so a point covers every pixel the sigmoid does not underflow on; the softmax runs over
    ALL points with uncovered ones scored 0 (``z * mask``, not a masked fill).  Returns (M [+1], P)."""
    dtype = v.dtype
    eps = torch.finfo(dtype).eps
    yy, xx = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    pix = torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=1).to(dtype)          # rasterer.py:25-27
    p2 = project_pixels(K, v, (width, height))
    diff = p2.view(-1, 1, 2) - pix.unsqueeze(0)
    radius = (K[0, 0] * diam / (v[:, 2] + eps)).abs().unsqueeze(-1)
    cover = (torch.sigmoid((radius - diff.pow(2).sum(-1).sqrt()) * soft) > 0).detach().to(dtype)
    z = _depth_scores(v, gain)
    if add_bg:
        z = torch.cat([z, (z.min() - 1).view(1, 1)])
        cover = torch.cat([cover, torch.ones_like(cover[:1])])
    return torch.softmax(z * cover, dim=0) * cover


def circle_opt_weights(K, v, add_bg=True, diam=0.025, gain=10000.0, soft=5.0):
    """``inside_circle_opt`` (primitives.py:71-162, rasterer.py:97-100).  Each point stamps its 15 x 15 pixel
    neighbourhood (``Rasterer.grid_prim``, offsets -7..7) around its TRUNCATED pixel position, indices clamped
    to the image, duplicates summed by the sparse -> dense conversion; the sigmoid never underflows inside that
    window, so the mask is the clamped window itself.  The image size is taken from the principal point
    (``x_px = int(K[0,2]) * 2``), as the reference does.  Returns (M [+1], P)."""
    dtype = v.dtype
    x_px, y_px = int(K[0, 2].int().item()) * 2, int(K[1, 2].int().item()) * 2
    p2 = project_pixels(K, v, (x_px, y_px))
    oy, ox = torch.meshgrid(torch.arange(-7, 8), torch.arange(-7, 8), indexing="ij")
    offs = torch.stack([ox.reshape(-1), oy.reshape(-1)], dim=1)                    # rasterer.py:30-32
    idx = (offs.to(dtype).unsqueeze(0) + p2.unsqueeze(1)).long()                   # truncation toward zero
    idx = torch.max(torch.min(idx, torch.tensor([[x_px - 1, y_px - 1]])), torch.tensor([[0, 0]]))
    cover = torch.zeros(v.shape[0], y_px * x_px, dtype=dtype)
    flat = idx[..., 1] * x_px + idx[..., 0]
    cover.scatter_(1, flat, torch.ones_like(flat, dtype=dtype))
    z = _depth_scores(v, gain)
    if add_bg:
        z = torch.cat([z, (z.min() - 1).view(1, 1)])
        cover = torch.cat([cover, torch.ones(1, y_px * x_px, dtype=dtype)])
    score = z.expand(-1, cover.shape[1]).masked_fill(cover == 0, torch.finfo(dtype).min)
    return torch.softmax(score, dim=0) * cover


def disc_weights_bg(rays, v, m, radius=DISC_RADIUS, gain=DEPTH_GAIN):
    """``inside_surfel`` with ``add_bg=True`` (primitives.py:232-237): one extra row that covers every pixel, scored
    below the farthest surfel."""
    dtype = v.dtype
    eps = torch.finfo(dtype).eps
    a = (m * v).sum(1, keepdim=True)
    b = m @ rays.t()
    b = torch.where(b.abs() < RAY_CUTOFF, torch.full_like(b, eps).detach(), b)
    z = a / b
    hit = rays.unsqueeze(0) * z.unsqueeze(-1)
    gap = radius - (v.unsqueeze(1) - hit).pow(2).sum(-1).sqrt()
    inside = (gap.clamp(min=0) > 0).detach()
    zeta = -z * inside.to(dtype)
    nu = zeta.norm(dim=0).detach()
    score = (zeta / (nu.unsqueeze(0) + eps) + 1).clamp(min=0) * gain
    bg = ((-v[:, 2:] * gain).min() - 1).expand(1, score.shape[1])
    score = torch.cat([score, bg])
    inside = torch.cat([inside, torch.ones_like(inside[:1])])
    score = score.masked_fill(~inside, torch.finfo(dtype).min)
    return torch.softmax(score, dim=0) * inside.to(dtype)


def compose_bg(w, c, bg):
    """rasterer.py:107-126 with a background image ``bg`` (3,H,W): only ``color`` and ``mask`` can be composed (the
    reference's depth / normals lines mix M+1 weights with M values and fail to broadcast)."""
    col = torch.cat([((c + 1) / 2).unsqueeze(-1).expand(-1, -1, w.shape[1]), bg.reshape(1, 3, -1)])
    color = (w.unsqueeze(1) * col).sum(0).clamp(max=1)
    mask = w.sum(0, keepdim=True).clamp(max=1)
    return color, mask


# ----------------------------------------------------------------------------
# Losses  (pipelines/optimizer.py:166-198 and 200-237)
# ----------------------------------------------------------------------------
def nearest_neighbour(query: torch.Tensor, ref: torch.Tensor):
    """Exact 1-NN in float64 on the float32 data, as sklearn's KDTree does
    (optimizer.py:180-181).  Returns (dist float64 (Q,), idx (Q,))."""
    q = query.detach().double()
    r = ref.detach().double()
    best_d = torch.full((q.shape[0],), float("inf"), dtype=torch.float64)
    best_i = torch.zeros(q.shape[0], dtype=torch.long)
    for s in range(0, r.shape[0], 4096):
        d2 = ((q.unsqueeze(1) - r[s:s + 4096].unsqueeze(0)) ** 2).sum(-1)
        d, i = d2.min(dim=1)
        upd = d < best_d
        best_d = torch.where(upd, d, best_d)
        best_i = torch.where(upd, i + s, best_i)
    return best_d.sqrt(), best_i


def loss_3d(xyzf: torch.Tensor, lidar_scaled: torch.Tensor, scale_value: float, radius=NN_RADIUS):
    """mean_i || L_nn(i) - v_i ||_2 over pairs closer than radius/scale."""
    if xyzf.numel() == 0 or lidar_scaled.numel() == 0:
        return xyzf.new_zeros(())
    dist, idx = nearest_neighbour(xyzf, lidar_scaled)
    close = dist < radius / scale_value
    if int(close.sum()) == 0:
        return xyzf.new_zeros(())
    return (lidar_scaled[idx[close]] - xyzf[close]).norm(dim=1).mean()


def resize_nearest(img: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """F.interpolate(mode='nearest') (optimizer.py:135-137): src = floor(dst*in/out)."""
    _, ih, iw = img.shape
    ys = torch.clamp((torch.arange(height, dtype=torch.float32) * (ih / height)).floor().long(), max=ih - 1)
    xs = torch.clamp((torch.arange(width, dtype=torch.float32) * (iw / width)).floor().long(), max=iw - 1)
    return img[:, ys][:, :, xs]


def loss_2d(color: torch.Tensor, target: torch.Tensor, radius=WIN_RADIUS, thr=NOCS_THR, dense=False):
    """optimizer.py:200-237.  For every rendered pixel m (sum_c colour != 0):
    delta_m = min_{(h,w)} || T(:,h,w) * max(radius - |(h,w)-m|, 0) - colour(:,m) ||;
    loss = mean{delta_m < thr}.  ``dense=False`` searches the 9x9 window where
    the weight can be non-zero plus the shared 'weight 0' candidate ||colour_m||;
    ``dense=True`` is the reference's O(M*H*W) form (small images only)."""
    _, H, W = color.shape
    nz = color.sum(0).nonzero()                  # (M,2) as (h,w)
    if int(nz.sum()) == 0:                       # optimizer.py:214 (quirk: empty OR only pixel (0,0))
        return color.new_zeros(())
    px = color[:, nz[:, 0], nz[:, 1]].t()        # (M,3)
    if dense:
        hh, ww = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        g = torch.stack([hh, ww], -1).to(color.dtype).reshape(1, -1, 2)
        wgt = (radius - (g - nz.view(-1, 1, 2).to(color.dtype)).pow(2).sum(-1).sqrt()).clamp(min=0)
        cand = target.reshape(1, 3, -1) * wgt.unsqueeze(1)
        best = (cand - px.unsqueeze(-1)).pow(2).sum(1).sqrt().min(dim=1)[0]
    else:
        r = int(math.ceil(radius)) - 1
        offs = torch.arange(-r, r + 1)
        dh, dw = torch.meshgrid(offs, offs, indexing="ij")
        dh, dw = dh.reshape(-1), dw.reshape(-1)
        h = nz[:, :1] + dh.unsqueeze(0)
        w = nz[:, 1:] + dw.unsqueeze(0)
        ok = (h >= 0) & (h < H) & (w >= 0) & (w < W)
        wgt = (radius - (dh.to(color.dtype) ** 2 + dw.to(color.dtype) ** 2).sqrt()).clamp(min=0)
        t = target[:, h.clamp(0, H - 1), w.clamp(0, W - 1)]               # (3,M,81)
        d = (t * wgt.view(1, 1, -1) - px.t().unsqueeze(-1)).pow(2).sum(0).sqrt()
        d = torch.where(ok, d, torch.full_like(d, float("inf")))
        best = d.min(dim=1)[0]
        # any pixel at distance >= radius contributes the candidate ||colour_m||
        corners = torch.tensor([[0, 0], [0, W - 1], [H - 1, 0], [H - 1, W - 1]], dtype=color.dtype)
        far = ((nz.to(color.dtype).unsqueeze(1) - corners.unsqueeze(0)).pow(2).sum(-1).sqrt().max(dim=1)[0]
               >= radius)
        zero_cand = px.pow(2).sum(1).sqrt()
        best = torch.where(far & (zero_cand < best), zero_cand, best)
    sel = best < thr
    return best[sel].mean()                       # NaN when empty, like the reference


# ----------------------------------------------------------------------------
# One refine iteration and the loop  (pipelines/optimizer.py:56-164)
# ----------------------------------------------------------------------------
@dataclass
class RefineState:
    yaw: torch.Tensor       # (1,)
    trans: torch.Tensor     # (3,)
    scale: torch.Tensor     # (1,)
    latent: torch.Tensor    # (L,)
    adam_m: Dict[str, torch.Tensor] = field(default_factory=dict)
    adam_v: Dict[str, torch.Tensor] = field(default_factory=dict)
    adam_t: int = 0

    @staticmethod
    def create(yaw, trans, scale, latent, dtype=torch.float32) -> "RefineState":
        mk = lambda a: torch.tensor(np.asarray(a, dtype=np.float32).reshape(-1)).to(dtype)
        return RefineState(mk(yaw), mk(trans), mk(scale), mk(latent))

    def as_numpy(self):
        return {k: getattr(self, k).detach().numpy().copy() for k in ("yaw", "trans", "scale", "latent")}


LR_ADAM = {"yaw": 0.01, "trans": 0.01}          # optimizer.py:34-36 (per-group lr wins over 0.03)
LR_SGD = {"scale": 0.01, "latent": 0.00003}     # optimizer.py:37-38


def iteration_losses(p: DecoderParams, pts, K, width, height, state: RefineState, target_full, lidar,
                     w2d: float, w3d: float, tile_rows=None, dense_2d=False):
    """Forward of one iteration with leaves (yaw, trans, scale, latent).  Returns
    a dict with both losses, the rendering and the surfels (graph attached)."""
    dt = pts.dtype
    lidar_s = torch.as_tensor(lidar, dtype=torch.float32).to(state.scale.device) / state.scale   # optimizer.py:84
    lidar_s = lidar_s.to(dt)
    pose = yaw_pose(state.yaw, state.trans)                                       # 87-90
    lat = torch.nn.functional.normalize(state.latent, p=2, dim=0)                 # 96
    sdf, nrm, _ = sdf_and_normals(p, lat, pts)                                    # 99-104
    sp, _, sn, keep = surface_points(pts, sdf, nrm)
    r = render(K, width, height, sp, sn, sn, pose, rot="dcm", output_nocs=True, tile_rows=tile_rows)
    out = {"render": r, "surf_pts": sp, "surf_nrm": sn, "keep": keep, "sdf": sdf, "skip": False}
    if r["xyzf"].numel() == 0 or lidar_s.numel() == 0:                            # 127-129
        out["skip"] = True
        return out
    l3 = loss_3d(r["xyzf"], lidar_s, float(state.scale[0]))
    tgt = resize_nearest(target_full.to(dt), height, width)
    l2 = loss_2d(r["color"], tgt, dense=dense_2d)
    out.update(loss_3d=l3, loss_2d=l2, loss=w3d * l3 + w2d * l2, target=tgt)
    if bool(torch.isnan(out["loss"])) or float(out["loss"]) == 0.0:               # 149-151
        out["skip"] = True
    return out


def refine_iteration(p, pts, K, width, height, state: RefineState, target_full, lidar, w2d, w3d,
                     tile_rows=None, dense_2d=False):
    """One pass of the loop body; mutates ``state`` (Adam on yaw/trans with
    betas (0.9,0.999), eps 1e-8; plain SGD on scale/latent).  Returns the
    forward dict plus the gradients."""
    leaves = {k: getattr(state, k).detach().clone().requires_grad_(True) for k in ("yaw", "trans", "scale", "latent")}
    tmp = RefineState(**leaves)
    out = iteration_losses(p, pts, K, width, height, tmp, target_full, lidar, w2d, w3d, tile_rows, dense_2d)
    if out["skip"]:
        out["grads"] = None
        return out
    extra = [t for t in list(p.weight) + list(p.bias) if t.requires_grad]   # reference-cost mode (optimizer.py:156)
    grads = torch.autograd.grad(out["loss"], list(leaves.values()) + extra, allow_unused=True)[:len(leaves)]
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(leaves, grads)}
    out["grads"] = grads
    state.adam_t += 1
    t = state.adam_t
    b1, b2, eps = 0.9, 0.999, 1e-8
    for k, lr in LR_ADAM.items():
        g = grads[k]
        m = state.adam_m.get(k, torch.zeros_like(g)) * b1 + (1 - b1) * g
        v = state.adam_v.get(k, torch.zeros_like(g)) * b2 + (1 - b2) * g * g
        state.adam_m[k], state.adam_v[k] = m, v
        denom = v.sqrt() / math.sqrt(1 - b2 ** t) + eps
        setattr(state, k, (getattr(state, k) - (lr / (1 - b1 ** t)) * m / denom).detach())
    for k, lr in LR_SGD.items():
        setattr(state, k, (getattr(state, k) - lr * grads[k]).detach())
    return out


def refine(p, pts, K, width, height, state: RefineState, target_full, lidar, w2d, w3d, iters,
           tile_rows=None, trace: Optional[list] = None):
    for _ in range(iters):
        out = refine_iteration(p, pts, K, width, height, state, target_full, lidar, w2d, w3d, tile_rows)
        if trace is not None:
            trace.append({
                "skip": out["skip"],
                "loss_2d": None if out["skip"] and "loss_2d" not in out else float(out["loss_2d"]),
                "loss_3d": None if out["skip"] and "loss_3d" not in out else float(out["loss_3d"]),
                **{k: v.copy() for k, v in state.as_numpy().items()},
            })
    return state
