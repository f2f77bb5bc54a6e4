"""Initial-pose RANSAC: device path (utils.pose.PoseEstimator) against the CPU port of the reference
(oracle with sklearn KD-trees, as utils/pose.py builds them) on the same seeded inputs.
    python tools/pose_bench.py            # prints a markdown table"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pose_oracle as PO            # dev tool: the CPU leg is the oracle
from sdflabel_b200 import _lib
from sdflabel_b200.utils.pose import PoseEstimator, nn_query, ransac_score

dev = torch.device("cuda")
print("| scene points | model points | CPU port (KD-trees) ms | device path ms | of which the two kernels ms | same pose |")
print("|---|---|---|---|---|---|")
for n_scene, n_out, n_model in ((300, 60, 1500), (700, 150, 3000), (1600, 400, 6000)):
    sc = PO.make_pose_scene(seed=21, n_model=n_model, n_scene=n_scene, n_outliers=n_out)
    args = [sc[k] for k in ("model_pts", "model_cls", "scene_pts", "scene_cls")]
    np.random.seed(5)
    t0 = time.perf_counter()
    ref = PO.init_pose_3d(*args, type="kabsch", scale_model=2.0, nn="kdtree")
    cpu_ms = (time.perf_counter() - t0) * 1e3
    targs = [torch.from_numpy(a).to(dev) for a in args]
    ts = []
    for rep in range(4):
        np.random.seed(5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = PoseEstimator.init_pose_3d(*targs, type="kabsch", scale_model=2.0)
        ts.append((time.perf_counter() - t0) * 1e3)
    # kernels alone: all 567 hypotheses scored
    T = torch.eye(3, 4, device=dev).repeat(567, 1, 1).contiguous()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nn_query(targs[3], targs[1]); ransac_score(targs[2], targs[3], targs[0], targs[1], T, 0.15, 0.15)
    a.record()
    nn_query(targs[3], targs[1]); ransac_score(targs[2], targs[3], targs[0], targs[1], T, 0.15, 0.15)
    b.record(); torch.cuda.synchronize()
    same = ref is not None and got is not None and np.array_equal(ref["rot"], got["rot"]) and np.array_equal(ref["tra"], got["tra"])
    print(f"| {n_scene + n_out} | {n_model} | {cpu_ms:.0f} | {min(ts):.1f} | {a.elapsed_time(b):.2f} | {same} |", flush=True)
