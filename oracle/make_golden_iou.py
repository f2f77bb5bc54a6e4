"""Writes tests/golden/rotate_iou.npz by running the UNMODIFIED reference kernel
``pipelines/rotate_iou.py::rotate_iou_gpu_eval`` under numba's CUDA simulator (TEST INFRASTRUCTURE;
needs /root/reference, build container only):

    NUMBA_ENABLE_CUDASIM=1 python -m oracle.make_golden_iou

The reference module selects a CUDA device through mpi4py at import (rotate_iou.py:10-14); both are
stubbed here (one rank, one simulated device).  The simulator executes the kernel's Python source with
numpy float32 scalars: same algorithm and thresholds, rounding that can differ in the last bits from the
compiled kernel (which contracts multiply-adds) - hence the 2e-5 tolerance of the parity tests.
Also stores the two host-side helpers of the same file (d3_box_overlap_kernel, image_box_overlap).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_boxes(rng, n, spread=4.0):
    """[x, y, w, h, angle] rows."""
    return np.concatenate([rng.uniform(-spread, spread, (n, 2)), rng.uniform(0.8, 4.5, (n, 2)),
                           rng.uniform(-3.3, 3.3, (n, 1))], 1).astype(np.float32)


def special_boxes():
    b = np.array([
        [0.0, 0.0, 2.0, 4.0, 0.0],        # reference box
        [0.0, 0.0, 2.0, 4.0, 0.0],        # identical
        [0.5, 0.25, 2.0, 4.0, 0.0],       # axis-aligned shift
        [0.0, 0.0, 2.0, 4.0, 1.5707964],  # rotated by 90 degrees
        [0.0, 0.0, 1.0, 1.0, 0.3],        # contained
        [10.0, 10.0, 2.0, 4.0, 0.7],      # disjoint
        [2.0, 0.0, 2.0, 4.0, 0.0],        # touching edge
        [0.3, -0.2, 3.0, 1.5, 0.785398],  # 45 degrees
        [1.0, 2.0, 2.0, 4.0, -0.4],
    ], dtype=np.float32)
    # (a copy rotated by pi is left out: its corners coincide up to rounding, the reference then finds more
    #  than 8 candidate vertices and writes past its 16-float local array - rotate_iou.py:190-199)
    return b


def main():
    if os.environ.get("NUMBA_ENABLE_CUDASIM") != "1":
        raise SystemExit("run with NUMBA_ENABLE_CUDASIM=1")
    import torch
    m = types.ModuleType("mpi4py")
    m.MPI = types.SimpleNamespace(COMM_WORLD=types.SimpleNamespace(Get_rank=lambda: 0))
    sys.modules["mpi4py"] = m
    torch.cuda.device_count = lambda: 1
    sys.path.insert(0, os.environ.get("SDFLABEL_REFERENCE", "/root/reference"))
    import pipelines.rotate_iou as R      # the reference's

    rng = np.random.RandomState(17)
    out = {}
    boxes, query = random_boxes(rng, 22), random_boxes(rng, 19)
    out["boxes"], out["query"] = boxes, query
    for crit in (-1, 0, 1, 2):
        out[f"iou_crit{crit}"] = R.rotate_iou_gpu_eval(boxes, query, crit)
    sp = special_boxes()
    out["special"] = sp
    out["special_iou"] = R.rotate_iou_gpu_eval(sp, sp, -1)
    out["special_inter"] = R.rotate_iou_gpu_eval(sp, sp, 2)
    # ragged sizes around the kernel's 64-box tiles
    b2, q2 = random_boxes(rng, 66, 6.0), random_boxes(rng, 3, 6.0)
    out["boxes_ragged"], out["query_ragged"] = b2, q2
    out["iou_ragged"] = R.rotate_iou_gpu_eval(b2, q2, -1)
    # 3D overlap (camera coordinates) as detection_3d.py:666-668 composes it
    b3 = np.concatenate([rng.uniform(-5, 5, (9, 3)), rng.uniform(1.0, 4.0, (9, 3)), rng.uniform(-3, 3, (9, 1))], 1).astype(np.float32)
    q3 = np.concatenate([rng.uniform(-5, 5, (7, 3)), rng.uniform(1.0, 4.0, (7, 3)), rng.uniform(-3, 3, (7, 1))], 1).astype(np.float32)
    q3[:4, :3] = b3[:4, :3] + rng.normal(scale=0.3, size=(4, 3)).astype(np.float32)
    out["boxes3d"], out["query3d"] = b3, q3
    for crit in (-1, 0, 1):
        rinc = R.rotate_iou_gpu_eval(b3[:, [0, 2, 3, 5, 6]], q3[:, [0, 2, 3, 5, 6]], 2)
        R.d3_box_overlap_kernel(b3, q3, rinc, crit, True)
        out[f"d3_camera_crit{crit}"] = rinc
    rinc = R.rotate_iou_gpu_eval(b3[:, [0, 1, 3, 4, 6]], q3[:, [0, 1, 3, 4, 6]], 2)
    R.d3_box_overlap_kernel(b3, q3, rinc, -1, False)
    out["d3_lidar"] = rinc
    ib = np.sort(rng.uniform(0, 100, (12, 2, 2)), axis=1).transpose(0, 2, 1).reshape(12, 4)   # [x1, y1, x2, y2]
    iq = np.sort(rng.uniform(0, 100, (8, 2, 2)), axis=1).transpose(0, 2, 1).reshape(8, 4)
    out["image_boxes"], out["image_query"] = ib, iq
    for crit in (-1, 0, 1, 2):
        out[f"image_crit{crit}"] = R.image_box_overlap(ib, iq, crit)
    np.savez_compressed(os.path.join(GOLDEN, "rotate_iou.npz"), **out)
    print("wrote", os.path.join(GOLDEN, "rotate_iou.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
