import cProfile, pstats, sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.grid import Grid3D
from sdflabel_b200.pipelines import optimizer as OPT
from sdflabel_b200.pipelines.optimizer import Optimizer
OPT.TEMPORAL_PRUNING = False
dev = torch.device("cuda")
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
grid = Grid3D(40, device=dev)
K = torch.from_numpy(sc["K"]); nocs = torch.from_numpy(sc["nocs_pred"]).pin_memory()
params = {k: v.copy() for k, v in sc["init"].items()}
opt = Optimizer(params, dev, sc["weights"])
for _ in range(5):
    opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"], viz_type=None)
torch.cuda.synchronize()
ts = []
for _ in range(200):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"], viz_type=None)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("e2e median ms", np.median(ts) * 1e3, "mean", np.mean(ts) * 1e3, "min", np.min(ts) * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"], viz_type=None)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
