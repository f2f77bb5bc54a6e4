"""CUDA pose RANSAC (sdfr_nn_query / sdfr_ransac_score behind utils.pose.PoseEstimator) against the
unmodified reference's results (tests/golden/pose_ransac.npz) and the oracle's per-hypothesis trace."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import pose_oracle as PO

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_ransac.npz")


@pytest.mark.parametrize("name", ["kabsch", "procrustes", "kabsch_large", "no_consensus"])
def test_pose_matches_reference(name):
    from sdflabel_b200.utils.pose import PoseEstimator
    g = np.load(GOLDEN)
    typ, scale, seed = str(g[f"{name}/type"]), float(g[f"{name}/scale_model"]), int(g[f"{name}/seed"])
    args = [torch.from_numpy(g[f"{name}/{k}"]).cuda() for k in ("model_pts", "model_cls", "scene_pts", "scene_cls")]
    model_before = args[0].clone()
    np.random.seed(seed)
    t0 = time.perf_counter()
    pose = PoseEstimator.init_pose_3d(*args, type=typ, scale_model=scale)
    dt = time.perf_counter() - t0
    assert np.random.randint(0, 2 ** 31 - 1) == int(g[f"{name}/rng_after"])     # same RNG consumption
    assert torch.equal(args[0], model_before)                                    # caller's cloud untouched
    assert (pose is not None) == bool(g[f"{name}/found"])
    if pose is not None:
        # same hypotheses, same inlier set -> the final fit is the reference's, bit for bit
        assert np.array_equal(np.asarray(pose["rot"], np.float64), g[f"{name}/rot"])
        assert np.array_equal(np.asarray(pose["tra"], np.float64), g[f"{name}/tra"])
        assert float(pose["scale"]) == float(g[f"{name}/scale"])
    print(f"{name}: {dt * 1e3:.1f} ms")


def test_kernels_match_oracle_trace():
    """Per-hypothesis inlier counts and the winning mask, and the NN kernel, against the oracle."""
    from sdflabel_b200.utils import pose as prod
    g = np.load(GOLDEN)
    name = "kabsch"
    mp, mc, sp, sc = (g[f"{name}/{k}"] for k in ("model_pts", "model_cls", "scene_pts", "scene_cls"))
    scale = float(g[f"{name}/scale_model"])
    np.random.seed(int(g[f"{name}/seed"]))
    _, tr = PO.init_pose_3d(mp, mc, sp, sc, type="kabsch", scale_model=scale, return_trace=True)
    dev = torch.device("cuda")
    mp_s = torch.from_numpy(mp).to(dev) * scale
    d, i = prod.nn_query(torch.from_numpy(sc).to(dev), torch.from_numpy(mc).to(dev))
    d0, i0 = PO.nn_exact(sc, mc)
    assert np.array_equal(i.cpu().numpy(), i0) and np.array_equal(d.cpu().numpy(), d0)
    valid = np.nonzero(tr["valid"])[0]
    t_d = torch.from_numpy(tr["transforms"][valid].reshape(-1, 3, 4)).to(dev)
    counts, masks = prod.ransac_score(torch.from_numpy(sp).to(dev), torch.from_numpy(sc).to(dev), mp_s,
                                      torch.from_numpy(mc).to(dev), t_d, 0.15, np.float32(0.15))
    assert np.array_equal(counts.cpu().numpy(), tr["counts"][valid])
    best = int(np.argmax(tr["counts"][valid]))
    assert np.array_equal(np.nonzero(masks[best].cpu().numpy())[0], tr["best"])


def test_edge_cases():
    from sdflabel_b200.utils import pose as prod
    from sdflabel_b200.utils.pose import PoseEstimator
    sc = PO.make_pose_scene(seed=1, n_scene=3, n_outliers=1)
    assert PoseEstimator.init_pose_3d(sc["model_pts"], sc["model_cls"], sc["scene_pts"], sc["scene_cls"],
                                      type="kabsch") is None                    # fewer than 5 scene points
    dev = torch.device("cuda")
    # ragged sizes around the 256-thread block and the 1024-point staging chunk; duplicate reference points
    rng = np.random.RandomState(5)
    for q, m in ((1, 1), (255, 1023), (257, 1025), (600, 3000)):
        qs = rng.normal(size=(q, 3)).astype(np.float32)
        rs = rng.normal(size=(m, 3)).astype(np.float32)
        if m > 10:
            rs[7] = rs[3]                                                       # a tie: the lowest index wins
        d, i = prod.nn_query(torch.from_numpy(qs).to(dev), torch.from_numpy(rs).to(dev))
        d0, i0 = PO.nn_exact(qs, rs)
        assert np.array_equal(i.cpu().numpy(), i0) and np.array_equal(d.cpu().numpy(), d0)
    est = PoseEstimator("kabsch", 2.0)
    sc = PO.make_pose_scene(seed=9)
    np.random.seed(1)
    pose = est.estimate(sc["model_pts"], sc["model_cls"], sc["scene_pts"], sc["scene_cls"], None, None)
    assert pose is not None and np.abs(pose["rot"] - sc["rot"]).max() < 2e-2 and np.abs(pose["tra"] - sc["tra"]).max() < 2e-2
