mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -q --timeout 120 -x 2>&1 | tail -30 | grep -v Warn
timeout 300 python - <<'PY' 2>&1 | grep -v Warn
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import numpy as np, torch
import bench
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.renderer.tracer import SphereTracer
from oracle import sdf_oracle as O
dev = torch.device('cuda')
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
lat = torch.tensor(sc['init']['latent'], device=dev)
pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0])).to(dev)
for size in (64, 256, 512, 1024):
    K = torch.from_numpy(sc['K']).clone(); K[:2] *= size / 256.0
    tr = SphereTracer(K, (size, size)).to(dev)
    with torch.no_grad():
        r = tr(dec, lat, pose)
        torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = tr(dec, lat, pose); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(size, 'hits', int(r['mask'].sum().item()), 'fwd ms', np.median(ts[2:]), 'rays/s', size*size/np.median(ts[2:])*1e3)
PY
