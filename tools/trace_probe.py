"""Dev tool: one forward of trace mode per resolution (for an ncu launch list) and timings."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.renderer.tracer import SphereTracer
from oracle import scenes, sdf_oracle as O   # pose / intrinsics helpers only

dev = torch.device("cuda")
dec, L = setup_dsdf("assets/deepsdf_synth.pt", precision=torch.float32)
dec = dec.to(dev)
lat = torch.nn.functional.normalize(torch.tensor([0.6, 0.6, 0.5]), dim=0).to(dev)
pose = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0])).to(dev)
stats = "--stats" in sys.argv          # per-launch ray counts on stderr (sdfr_trace_set_stats(2): synchronises every launch)
sizes = [int(a) for a in sys.argv[1:] if a != "--stats"] or [64, 256]
if stats:
    from sdflabel_b200 import _lib
    _lib.load().sdfr_trace_set_stats(2)
for size in sizes:
    K = scenes.intrinsics(size)
    tracer = SphereTracer(K, (size, size)).to(dev)
    with torch.no_grad():
        for _ in range(3):
            r = tracer(dec, lat, pose)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            r = tracer(dec, lat, pose)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"trace {size}x{size}: hits {int(r['mask'].sum())} fwd {ms:.3f} ms {size*size/ms*1e3:,.0f} rays/s", flush=True)
