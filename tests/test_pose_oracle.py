"""The pose-RANSAC oracle against what the unmodified reference returned (tests/golden/pose_ransac.npz,
written by oracle/make_golden_pose.py), and the host-side pieces of the product mirror."""
import os

import numpy as np
import pytest

from oracle import pose_oracle as PO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pose_ransac.npz")


def _case(g, n):
    return (g[f"{n}/model_pts"], g[f"{n}/model_cls"], g[f"{n}/scene_pts"], g[f"{n}/scene_cls"],
            str(g[f"{n}/type"]), float(g[f"{n}/scale_model"]), int(g[f"{n}/seed"]))


@pytest.mark.parametrize("name", ["kabsch", "procrustes", "no_consensus"])
def test_oracle_matches_reference(name):
    g = np.load(GOLDEN)
    mp, mc, sp, sc, typ, scale, seed = _case(g, name)
    np.random.seed(seed)
    pose = PO.init_pose_3d(mp, mc, sp, sc, type=typ, scale_model=scale)
    # the reference consumed exactly as much of the numpy RNG stream
    assert np.random.randint(0, 2 ** 31 - 1) == int(g[f"{name}/rng_after"])
    assert (pose is not None) == bool(g[f"{name}/found"])
    if pose is not None:
        assert np.array_equal(np.asarray(pose["rot"], np.float64), g[f"{name}/rot"])
        assert np.array_equal(np.asarray(pose["tra"], np.float64), g[f"{name}/tra"])
        assert float(pose["scale"]) == float(g[f"{name}/scale"])


def test_oracle_too_few_scene_points():
    sc = PO.make_pose_scene(seed=1, n_scene=3, n_outliers=1)
    assert PO.init_pose_3d(sc["model_pts"], sc["model_cls"], sc["scene_pts"], sc["scene_cls"], type="kabsch") is None


def test_nn_exact_is_a_nearest_neighbour_search():
    rng = np.random.RandomState(0)
    q, r = rng.normal(size=(50, 3)).astype(np.float32), rng.normal(size=(400, 3)).astype(np.float32)
    d, i = PO.nn_exact(q, r)
    full = np.linalg.norm(q[:, None].astype(np.float64) - r[None].astype(np.float64), axis=2)
    assert np.array_equal(i, full.argmin(1))
    assert np.allclose(d, full.min(1), rtol=1e-14)


def test_product_fits_equal_the_oracle_fits():
    """kabsch / procrustes of the product mirror (host numpy, same LAPACK calls) against the oracle's."""
    from sdflabel_b200.utils import pose as prod
    rng = np.random.RandomState(2)
    for _ in range(20):
        a = rng.normal(size=(4, 3)).astype(np.float32)
        b = rng.normal(size=(4, 3)).astype(np.float32)
        r0, t0 = PO.kabsch(a, b)
        r1, t1 = prod.kabsch(a, b)
        assert np.array_equal(r0, r1) and np.array_equal(t0, t1)
        p0, p1 = PO.procrustes(a, b), prod.procrustes(a, b)
        assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(p0, p1))
    assert prod.procrustes(np.zeros((4, 3), np.float32), np.zeros((4, 3), np.float32)) is None
    # the batched form is the same LAPACK / BLAS calls: same bits (reflections included)
    a = rng.normal(size=(300, 4, 3)).astype(np.float32)
    b = rng.normal(size=(300, 4, 3)).astype(np.float32)
    rots, tras = prod.kabsch_batch(a, b)
    for i in range(300):
        r0, t0 = PO.kabsch(a[i], b[i])
        assert np.array_equal(r0, rots[i]) and np.array_equal(t0, tras[i])


def test_product_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sdflabel_b200 import _lib
    from sdflabel_b200.utils.pose import PoseEstimator
    sc = PO.make_pose_scene(seed=1)
    with pytest.raises(_lib.SdfrError):
        PoseEstimator.init_pose_3d(sc["model_pts"], sc["model_cls"], sc["scene_pts"], sc["scene_cls"], type="kabsch")


def test_legacy_choice_consumes_numpy_stream_exactly():
    """sdfr_np_choice4 (host loop over numpy's MT19937 state) == the np.random.choice calls of pose.py:139,
    numbers and generator state alike."""
    from sdflabel_b200.utils.pose import legacy_choice4
    for seed, n in ((1, 5), (2, 360), (3, 851), (4, 4), (5, 70000)):
        draws = 567 if n < 10000 else 6
        np.random.seed(seed)
        np.random.rand(seed * 311)                       # somewhere inside the 624-word block
        ref = np.stack([np.random.choice(range(n), 4, replace=False) for _ in range(draws)])
        after_ref = np.random.randint(0, 2 ** 31 - 1, size=4)
        np.random.seed(seed)
        np.random.rand(seed * 311)
        got = legacy_choice4(n, draws)
        after = np.random.randint(0, 2 ** 31 - 1, size=4)
        assert np.array_equal(ref, got) and np.array_equal(after_ref, after), (seed, n)
    np.random.seed(9)
    np.random.normal(size=3)                             # a cached gaussian in the state must survive the round trip
    a = legacy_choice4(100, 3)
    x = np.random.normal()
    np.random.seed(9)
    np.random.normal(size=3)
    b = np.stack([np.random.choice(range(100), 4, replace=False) for _ in range(3)])
    assert np.array_equal(a, b) and x == np.random.normal()
