"""Dev tool (build container only): the UNMODIFIED reference's Optimizer.optimize beside the oracle port, same scene,
same host - how conservative is the `kind: "port"` CPU arm of bench.py?  cfg1 (64x64, D=40): the largest crop the
reference's M x P x 3 tensors hold on a small host."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness, prior as P, scenes, sdf_oracle as O
import bench

size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
threads = os.cpu_count() or 1
torch.set_num_threads(threads)
ref = ref_harness.load()
prior = P.load_prior(bench.PRIOR)
sc = scenes.make_scene(prior, size=size, density=bench.DENSITY)
dec, L = ref.setup_dsdf(bench.PRIOR, precision=torch.float32)
params = {k: v.copy() for k, v in sc["init"].items()}
opt = ref.Optimizer(params, torch.device("cpu"), sc["weights"])
ref.grid_module.grads.clear()
grid = ref.Grid3D(bench.DENSITY)
times = []
for i in range(4):
    t0 = time.perf_counter()
    opt.optimize(1, torch.tensor(sc["nocs_pred"]), sc["lidar"], dec, grid, torch.tensor(sc["K"]), sc["crop_size"], viz_type=None)
    times.append(time.perf_counter() - t0)
ref_s = float(np.mean(times[1:]))
port_s, _ = bench.cpu_iteration_time(P.load_prior(bench.PRIOR), sc, 3, 1, threads=threads, size=size)
print(json.dumps({"config": f"{size}x{size}, D={bench.DENSITY}, 1 detection, {threads} threads", "reference_unmodified_s_per_iteration": ref_s,
                  "oracle_port_s_per_iteration": port_s, "port_speedup_over_reference": ref_s / port_s}))
