"""Dev tool: where do the fused and the plain march disagree, and what does the field look like along those rays."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from sdflabel_b200 import _lib
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.renderer.tracer import SphereTracer
from oracle import prior as P, scenes, sdf_oracle as O, trace_oracle as T

dev = torch.device("cuda")
size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
yaw = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
dec, L = setup_dsdf("assets/deepsdf_synth.pt", precision=torch.float32); dec = dec.to(dev)
plain, _ = setup_dsdf("assets/deepsdf_synth.pt", precision=torch.float32); plain = plain.to(dev); plain.mlp_impl = _lib.MLP_FFMA; STEPS = 256
prior = P.load_prior("assets/deepsdf_synth.pt")
lat = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0)
pose = O.yaw_pose(torch.tensor([yaw]), torch.tensor([0.0, 0.0, 5.0]))
K = scenes.intrinsics(size)
tr = SphereTracer(K, (size, size)).to(dev)
with torch.no_grad():
    a = tr(dec, lat.to(dev), pose.to(dev), normalize_latent=False)
    b = SphereTracer(K, (size, size), max_steps=STEPS).to(dev)(plain, lat.to(dev), pose.to(dev), normalize_latent=False)
ma, mb = (a["mask"][0] > 0.5).cpu(), (b["mask"][0] > 0.5).cpu()
print("hits fused", int(ma.sum()), "plain", int(mb.sum()), "diff", int((ma ^ mb).sum()))
o, d, rn = T.rays(K, size, size, pose)
for (y, x) in (ma ^ mb).nonzero().tolist():
    j = y * size + x
    nb = mb[max(y-1,0):y+2, max(x-1,0):x+2]
    tau_a = float(a["depth"][0, y, x]) / float(rn[j, 2]); tau_b = float(b["depth"][0, y, x]) / float(rn[j, 2])
    ts = torch.linspace(3.5, 6.5, 3001)
    with torch.no_grad():
        f = O.decoder_forward(prior, torch.cat([lat.expand(ts.numel(), -1), o + ts[:, None] * d[j]], 1)).squeeze(1)
    neg = (f < 0).nonzero()
    first_neg = float(ts[neg[0, 0]]) if neg.numel() else None
    # local minima of f before the first negative sample
    k_end = int(neg[0, 0]) if neg.numel() else ts.numel()
    fm = f[:k_end]
    mins = [(round(float(ts[k]), 4), round(float(fm[k]), 5)) for k in range(1, k_end - 1) if fm[k] < fm[k-1] and fm[k] <= fm[k+1] and fm[k] < 0.02]
    print(f"px ({y},{x}) fused={bool(ma[y,x])} plain={bool(mb[y,x])} plain-neighbourhood hits {int(nb.sum())}/9 tau fused {tau_a:.4f} plain {tau_b:.4f} first f<0 at {first_neg} minima before: {mins[:4]}")
