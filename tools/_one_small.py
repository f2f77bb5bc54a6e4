import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.pipelines.optimizer import _engine_for
dev = torch.device("cuda")
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
nocs = torch.from_numpy(sc["nocs_pred"])
size = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = torch.from_numpy(sc["K"]).clone(); K[:2] *= size / 256.0
eng = _engine_for(dec, 1, 40, size, size, sc["lidar"].shape[0], 64, sc["weights"], dec.mlp_impl)
eng.set_active(1)
eng.set_detection(0, K, size, size, nocs, sc["lidar"], sc["init"]["yaw"], sc["init"]["trans"], sc["init"]["scale"], sc["init"]["latent"])
eng.run(4)
torch.cuda.synchronize()
m = int(eng.view(0, 'surf_count').item())
bb = eng.view(0, 'cam_pts', 3 * m)
print("surfels", m)
