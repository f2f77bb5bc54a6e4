"""Host-side profile of Optimizer.optimize(1) (where do the e2e milliseconds go?)."""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.grid import Grid3D
from sdflabel_b200.pipelines.optimizer import Optimizer
dev = torch.device("cuda")
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
grid = Grid3D(40, device=dev)
K = torch.from_numpy(sc["K"]); nocs = torch.from_numpy(sc["nocs_pred"]).pin_memory()
params = {k: v.copy() for k, v in sc["init"].items()}
opt = Optimizer(params, dev, sc["weights"])
for _ in range(3):
    opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"])
torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"]); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("optimize(1) wall ms: median %.3f min %.3f max %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3, max(ts) * 1e3))
ts = []
for _ in range(5):
    t0 = time.perf_counter(); opt.optimize(60, nocs, sc["lidar"], dec, grid, K, sc["crop_size"]); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("optimize(60) wall ms: median %.3f  -> %.3f ms/iter" % (np.median(ts) * 1e3, np.median(ts) * 1e3 / 60))
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:4000])
