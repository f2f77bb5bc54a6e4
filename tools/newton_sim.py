"""Design study (dev tool): convergence of the Newton finish from the hand-over band on the stock prior."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import prior as P, scenes, sdf_oracle as O, trace_oracle as T
torch.set_num_threads(16)
size = 128
prior = P.load_prior("assets/deepsdf_synth.pt")
K = scenes.intrinsics(size)
lat = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0)
pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0]))
o, d, rn = T.rays(K, size, size, pose)
inv = 1.0 / d
ta, tb = (T.BOX_LO - o) * inv, (T.BOX_HI - o) * inv
t0 = torch.minimum(ta, tb).max(dim=1)[0].clamp(min=0.0); t1 = torch.maximum(ta, tb).min(dim=1)[0]
def fG(idx, tau):
    x = (o + tau[idx, None] * d[idx]).requires_grad_(True)
    f = O.decoder_forward(prior, torch.cat([lat.expand(idx.numel(), -1), x], 1)).squeeze(1)
    (G,) = torch.autograd.grad(f.sum(), x)
    return f.detach(), G
near = float(sys.argv[1]) if len(sys.argv) > 1 else 5e-3
active = t0 <= t1; tau = t0.clone(); nearset = torch.zeros_like(active)
for step in range(64):
    idx = active.nonzero().squeeze(1)
    if not idx.numel(): break
    with torch.no_grad():
        f = O.decoder_forward(prior, torch.cat([lat.expand(idx.numel(), -1), o + tau[idx, None] * d[idx]], 1)).squeeze(1)
    nr = f.abs() < near
    nearset[idx[nr]] = True; active[idx[nr]] = False
    go = ~nr; tau[idx[go]] += f[go]
    out = (tau[idx] > t1[idx]) | (tau[idx] < 0); active[idx[out & go]] = False
idx = nearset.nonzero().squeeze(1)
print("near rays", idx.numel())
work = idx
for it in range(5):
    f, G = fG(work, tau)
    Gd = (G * d[work]).sum(1)
    q = torch.tensor([0.5, 0.9, 0.99, 1.0])
    print(f"round {it}: rows {work.numel()} |f| quantiles {[float(v) for v in f.abs().quantile(q)]}  |Gd| min {float(Gd.abs().min()):.3f}  converged(<5e-5) {int((f.abs()<5e-5).sum())}")
    conv = f.abs() < 5e-5
    step = torch.where(Gd.abs() > 1e-3, -f / Gd, f).clamp(-4 * near, 4 * near)
    tau[work[~conv]] += step[~conv]
    work = work[~conv]
    if not work.numel(): break
