"""Box overlaps of the evaluator: oracle and host helpers against the reference's results
(tests/golden/rotate_iou.npz, the unmodified numba kernel under the CUDA simulator), CUDA kernel
against both (-m gpu)."""
import os

import numpy as np
import pytest

from oracle import rotate_iou_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rotate_iou.npz")
# the compiled kernels (numba's and ours) contract multiply-adds and use float32 sin / cos, the simulator that
# wrote the goldens does neither; intersection vertices move by a few float32 ulps
TOL = 2e-5


def test_oracle_matches_reference():
    g = np.load(GOLDEN)
    for c in (-1, 0, 1, 2):
        assert np.array_equal(RO.rotate_iou(g["boxes"], g["query"], c), g[f"iou_crit{c}"])
    assert np.array_equal(RO.rotate_iou(g["special"], g["special"], -1), g["special_iou"])
    assert np.array_equal(RO.rotate_iou(g["special"], g["special"], 2), g["special_inter"])
    assert np.array_equal(RO.rotate_iou(g["boxes_ragged"], g["query_ragged"], -1), g["iou_ragged"])
    # sanity of the fixture itself: identical boxes overlap fully, disjoint ones not at all
    assert g["special_iou"][0, 1] == pytest.approx(1.0, abs=1e-6) and g["special_iou"][0, 5] == 0.0
    assert g["special_inter"][0, 4] == pytest.approx(1.0, abs=1e-5)          # the contained 1 x 1 box


def test_host_helpers_match_reference():
    from sdflabel_b200.pipelines import rotate_iou as prod
    g = np.load(GOLDEN)
    b3, q3 = g["boxes3d"], g["query3d"]
    for c in (-1, 0, 1):
        rinc = RO.rotate_iou(b3[:, [0, 2, 3, 5, 6]], q3[:, [0, 2, 3, 5, 6]], 2)
        want = g[f"d3_camera_crit{c}"]
        assert np.array_equal(RO.d3_box_overlap(b3, q3, rinc, c, True), want)
        prod.d3_box_overlap_kernel(b3, q3, rinc, c, True)                    # in place, like the reference
        assert np.allclose(rinc, want, rtol=1e-6, atol=1e-7)
    rinc = RO.rotate_iou(b3[:, [0, 1, 3, 4, 6]], q3[:, [0, 1, 3, 4, 6]], 2)
    prod.d3_box_overlap_kernel(b3, q3, rinc, -1, False)
    assert np.allclose(rinc, g["d3_lidar"], rtol=1e-6, atol=1e-7)
    for c in (-1, 0, 1, 2):
        want = g[f"image_crit{c}"]
        assert np.array_equal(RO.image_box_overlap(g["image_boxes"], g["image_query"], c), want)
        assert np.allclose(prod.image_box_overlap(g["image_boxes"], g["image_query"], c), want, rtol=1e-12)
    assert prod.div_up(65, 64) == 2 and prod.div_up(64, 64) == 1
    assert prod.rotate_iou_gpu_eval(np.zeros((0, 5), np.float32), np.zeros((3, 5), np.float32)).shape == (0, 3)


@pytest.mark.gpu
def test_cuda_kernel_vs_reference_and_oracle():
    from sdflabel_b200.pipelines import rotate_iou as prod
    g = np.load(GOLDEN)
    for c in (-1, 0, 1, 2):
        got = prod.rotate_iou_gpu_eval(g["boxes"], g["query"], c)
        assert got.dtype == np.float32 and np.abs(got - g[f"iou_crit{c}"]).max() < TOL * (10 if c == 2 else 1)
    assert np.abs(prod.rotate_iou_gpu_eval(g["special"], g["special"], -1) - g["special_iou"]).max() < TOL
    assert np.abs(prod.rotate_iou_gpu_eval(g["boxes_ragged"], g["query_ragged"], -1) - g["iou_ragged"]).max() < TOL
    # float64 boxes come back as float32: the reference casts `boxes` first, then `iou.astype(boxes.dtype)`
    assert prod.rotate_iou_gpu_eval(g["boxes"].astype(np.float64), g["query"].astype(np.float64)).dtype == np.float32
    # and the empty case keeps that dtype (rotate_iou.py:308-310)
    assert prod.rotate_iou_gpu_eval(np.zeros((0, 5)), g["query"]).dtype == np.float32
    # seeded larger case against the oracle: several 64 x 64 tiles, ragged edges
    rng = np.random.RandomState(3)
    from oracle.make_golden_iou import random_boxes
    b, q = random_boxes(rng, 150, 5.0), random_boxes(rng, 70, 5.0)
    want = RO.rotate_iou(b, q, -1)
    got = prod.rotate_iou_gpu_eval(b, q, -1)
    assert np.abs(got - want).max() < TOL and (got > 0).sum() == (want > 0).sum()
    # the evaluator's 3D composition (detection_3d.py:666-668)
    b3, q3 = g["boxes3d"], g["query3d"]
    rinc = prod.rotate_iou_gpu_eval(b3[:, [0, 2, 3, 5, 6]], q3[:, [0, 2, 3, 5, 6]], 2)
    prod.d3_box_overlap_kernel(b3, q3, rinc, -1, True)
    assert np.abs(rinc - g["d3_camera_crit-1"]).max() < TOL
