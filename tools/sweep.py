"""cfg3 / cfg5 sweeps: refine-step throughput over batch size and crop size
(splat mode, fused engine) and sphere-tracing forward / forward+backward throughput over resolution.
Writes a markdown table to stdout.  python tools/sweep.py > gpurun_out/sweep.md"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sdflabel_b200 import _lib
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.grid import Grid3D
from sdflabel_b200.pipelines.optimizer import _engine_for
from sdflabel_b200.renderer.tracer import SphereTracer

dev = torch.device("cuda")
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


print("## Splat mode: one refine iteration (fwd + bwd + update), fused engine, Grid3D(40), stock prior\n")
print("| detections / launch | crop | ms / step | detections*iterations / s | rays / s |")
print("|---|---|---|---|---|")
nocs = torch.from_numpy(sc["nocs_pred"])
for size in (64, 256):
    K = torch.from_numpy(sc["K"]).clone()
    K[:2] *= size / 256.0
    for B in (1, 4, 16, 32, 64):
        eng = _engine_for(dec, B, 40, size, size, sc["lidar"].shape[0], 64, sc["weights"], dec.mlp_impl)
        eng.set_active(B)
        for b in range(B):
            eng.set_detection(b, K, size, size, nocs, sc["lidar"], sc["init"]["yaw"], sc["init"]["trans"],
                              sc["init"]["scale"], sc["init"]["latent"])
        ms = timed(lambda: eng.run(1))
        print(f"| {B} | {size}x{size} | {ms:.3f} | {B / ms * 1e3:,.0f} | {B * size * size / ms * 1e3:,.0f} |", flush=True)
        eng.get(0)

print("\n## Trace mode (cfg5): sphere tracing of one latent (eps 1e-4), stock prior; batch = poses rendered per call "
      "(`SphereTracer.render_views`, 4 / 8 CUDA streams); tensor fraction = algorithmic decoder flops of the evaluated rows / "
      "time / measured dense bf16 peak\n")
print("| resolution | hit rays / view | batch | streams | fwd ms / call | fwd rays / s | tensor frac | fwd+bwd ms (batch 1) |")
print("|---|---|---|---|---|---|---|---|")
import ctypes as C, json
lib = _lib.load()
spec = json.load(open(os.path.splitext(bench.PRIOR)[0] + ".json"))
flop_pt, pk = bench.mlp_flops_per_point(spec), bench.peaks()
lat = torch.tensor(sc["init"]["latent"], device=dev)
from oracle import sdf_oracle as O   # pose helper only (test infrastructure; this script is a dev tool)
pose_h = O.yaw_pose(torch.tensor([0.6]), torch.tensor([0.0, 0.0, 5.0]))
pose = pose_h.to(dev)
for size in (64, 128, 256, 512, 1024):
    K = torch.from_numpy(sc["K"]).clone()
    K[:2] *= size / 256.0
    tracer = SphereTracer(K, (size, size)).to(dev)
    with torch.no_grad():
        r = tracer(dec, lat, pose)
        hits = int(r["mask"].sum().item())
        fwd = timed(lambda: tracer(dec, lat, pose), n=5, warm=2)
        lib.sdfr_trace_set_stats(1)
        tracer(dec, lat, pose)
        torch.cuda.synchronize()
        cnt = (C.c_int64 * 4)()
        lib.sdfr_trace_get_stats(cnt)
        lib.sdfr_trace_set_stats(0)
    flops = flop_pt * (cnt[1] + 2.0 * cnt[3])          # march rows + Newton rows (forward + input gradient), cache reused

    def fb():
        l = lat.clone().requires_grad_(True); p = pose.clone().requires_grad_(True)
        out = tracer(dec, l, p)
        (out["depth"].sum() + out["color"].sum()).backward()
    both = timed(fb, n=5, warm=2)
    print(f"| {size}x{size} | {hits} | 1 | 1 | {fwd:.2f} | {size * size / fwd * 1e3:,.0f} | {flops / (fwd * 1e-3) / 1e12 / pk['tensor']:.2f} | {both:.2f} |", flush=True)
    for batch in ((4, 16, 64, 128) if size <= 256 else (4, 16)):
        views = [O.yaw_pose(torch.tensor([0.6 + 0.05 * i]), torch.tensor([0.0, 0.0, 5.0])) for i in range(batch)]
        for streams in (4, 8):
            if streams > batch:
                continue
            ms = timed(lambda: tracer.render_views(dec, lat, views, views_in_flight=streams), n=3, warm=1)
            print(f"| {size}x{size} | {hits} | {batch} | {streams} | {ms:.2f} | {batch * size * size / ms * 1e3:,.0f} | {batch * flops / (ms * 1e-3) / 1e12 / pk['tensor']:.2f} | |", flush=True)
