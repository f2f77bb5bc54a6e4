SDFR_LIB=sdflabel_b200/libsdfr_dbg.so timeout 300 python - <<'PY' 2>&1 | grep "march launch\|size" 
import sys, torch
sys.path.insert(0, ".")
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.renderer.tracer import SphereTracer
from oracle import scenes, sdf_oracle as O
dev = torch.device("cuda")
dec, L = setup_dsdf("assets/deepsdf_synth.pt", precision=torch.float32); dec = dec.to(dev)
lat = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0).to(dev)
pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0])).to(dev)
for size in (64, 256):
    print("size", size, flush=True)
    tr = SphereTracer(scenes.intrinsics(size), (size, size)).to(dev)
    with torch.no_grad():
        r = tr(dec, lat, pose, normalize_latent=False)
    torch.cuda.synchronize()
PY
