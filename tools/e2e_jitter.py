import gc, os, sys, time
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
opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"])
eng = opt.engine
def stats(name, ts):
    ts = np.array(ts)
    print(f"{name:28s} n={len(ts)} median {np.median(ts):.3f} ms  mean {ts.mean():.3f}  >10ms: {int((ts>10).sum())}  max {ts.max():.1f}  idx {np.nonzero(ts>10)[0].tolist()}")
x = torch.zeros(1 << 20, device=dev)
ts = []
for _ in range(300):
    t0 = time.perf_counter(); x.add_(1); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
stats("torch add_+sync", ts)
ts = []
for _ in range(300):
    t0 = time.perf_counter(); x.add_(1); y = x[:4].cpu(); ts.append((time.perf_counter() - t0) * 1e3)
stats("torch add_+.cpu()", ts)
ts = []
for _ in range(100):
    t0 = time.perf_counter(); eng.run(1); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
stats("eng.run(1)+sync", ts)
ts = []
for _ in range(100):
    t0 = time.perf_counter(); eng.run(1); eng.get(0); ts.append((time.perf_counter() - t0) * 1e3)
stats("eng.run(1)+get", ts)
ts = []
for _ in range(100):
    t0 = time.perf_counter(); opt.optimize(1, nocs, sc["lidar"], dec, grid, K, sc["crop_size"]); ts.append((time.perf_counter() - t0) * 1e3)
stats("optimize(1)", ts)
