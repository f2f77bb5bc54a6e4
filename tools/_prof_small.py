"""Per-stage times of one refine iteration at small crops (B = 1 and 32), pruned engine."""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sdflabel_b200 import _lib
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.pipelines.optimizer import _engine_for
lib = _lib.load()
dev = torch.device("cuda")
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
nocs = torch.from_numpy(sc["nocs_pred"])
for size in (32, 64, 256):
    K = torch.from_numpy(sc["K"]).clone(); K[:2] *= size / 256.0
    for B in (1, 32):
        eng = _engine_for(dec, B, 40, size, size, sc["lidar"].shape[0], 64, sc["weights"], dec.mlp_impl)
        eng.set_active(B)
        for b in range(B):
            eng.set_detection(b, K, size, size, nocs, sc["lidar"], sc["init"]["yaw"], sc["init"]["trans"], sc["init"]["scale"], sc["init"]["latent"])
        eng.run(3)
        n = C.c_int(0); ms = np.zeros(16, dtype=np.float32); rows = C.c_int32(0)
        _lib.check(lib.sdfr_refine_profile(eng.handle, 5, _lib.fptr(ms), 16, C.byref(n), C.byref(rows), _lib.stream_ptr()))
        print(size, B, {lib.sdfr_refine_stage_name(k).decode(): round(float(ms[k]) * 1e3, 1) for k in range(n.value)}, "rows", rows.value, flush=True)
