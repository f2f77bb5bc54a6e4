"""Box overlaps of the KITTI evaluator.

Mirror of the reference's pipelines/rotate_iou.py: ``rotate_iou_gpu_eval`` (289-325, numpy in / numpy
out, same ``criterion`` codes and argument order), ``d3_box_overlap_kernel`` (328-357, updates ``rinc``
in place) and ``image_box_overlap`` (360-379), as ``pipelines/detection_3d.py:31,557-668`` imports them.
The rotated overlap runs in ``sdfr_rotate_iou`` (csrc/rotate_iou.cu) instead of a numba.cuda kernel, so
importing this module neither needs numba nor initialises MPI / selects a CUDA device
(rotate_iou.py:10-14 does both at import time).  The two cheap host helpers are vectorised numpy.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def div_up(m, n):
    return m // n + (m % n > 0)


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """Rotated box overlap on the GPU.

    Args:
        boxes (np.ndarray [N, 5]): centres, dims, angles (clockwise when positive)
        query_boxes (np.ndarray [K, 5])
        criterion: -1 IoU, 0 intersection / query-box area, 1 intersection / box area, 2 intersection
        device_id: CUDA device
    Returns: np.ndarray [N, K], always float32: the reference rebinds ``boxes`` to its float32 copy before
        ``iou.astype(boxes.dtype)`` (rotate_iou.py:304-305,325), so float64 inputs come back as float32 too
    """
    n, k = boxes.shape[0], query_boxes.shape[0]
    if n == 0 or k == 0:
        return np.zeros((n, k), dtype=np.float32)
    if not torch.cuda.is_available():
        raise _lib.SdfrError("rotate_iou_gpu_eval needs a CUDA device (there is no CPU path)")
    dev = torch.device('cuda', int(device_id))
    with torch.cuda.device(dev):
        b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32)).to(dev)
        q = torch.from_numpy(np.ascontiguousarray(query_boxes, dtype=np.float32)).to(dev)
        iou = torch.empty((n, k), device=dev, dtype=torch.float32)
        _lib.check(_lib.load().sdfr_rotate_iou(b.data_ptr(), n, q.data_ptr(), k, int(criterion), iou.data_ptr(),
                                               _lib.stream_ptr()))
        return iou.cpu().numpy()


def d3_box_overlap_kernel(boxes, qboxes, rinc, criterion=-1, camera_coordinate=False):
    """BEV intersection areas ``rinc`` [N, K] -> 3D overlaps, in place (camera frame: y is the vertical
    axis and points down, boxes [x, y, z, l/w.., ry] as ``detection_3d.py:666-668`` passes them)."""
    b, q = np.asarray(boxes), np.asarray(qboxes)
    if camera_coordinate:
        iw = (np.minimum(b[:, None, 1], q[None, :, 1]) -
              np.maximum(b[:, None, 1] - b[:, None, 4], q[None, :, 1] - q[None, :, 4]))
    else:
        iw = (np.minimum(b[:, None, 2] + b[:, None, 5], q[None, :, 2] + q[None, :, 5]) -
              np.maximum(b[:, None, 2], q[None, :, 2]))
    area1 = (b[:, 3] * b[:, 4] * b[:, 5])[:, None]
    area2 = (q[:, 3] * q[:, 4] * q[:, 5])[None, :]
    inc = iw * rinc
    if criterion == -1:
        ua = area1 + area2 - inc
    elif criterion == 0:
        ua = np.broadcast_to(area1, inc.shape)
    elif criterion == 1:
        ua = np.broadcast_to(area2, inc.shape)
    else:
        ua = inc
    touched = rinc > 0
    with np.errstate(divide='ignore', invalid='ignore'):
        val = np.where(iw > 0, inc / ua, 0.0)
    rinc[touched] = val[touched].astype(rinc.dtype)


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """Axis-aligned 2D overlaps [N, K] of [x1, y1, x2, y2] boxes."""
    b, q = np.asarray(boxes), np.asarray(query_boxes)
    iw = np.minimum(b[:, None, 2], q[None, :, 2]) - np.maximum(b[:, None, 0], q[None, :, 0])
    ih = np.minimum(b[:, None, 3], q[None, :, 3]) - np.maximum(b[:, None, 1], q[None, :, 1])
    ba = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))[:, None]
    qa = ((q[:, 2] - q[:, 0]) * (q[:, 3] - q[:, 1]))[None, :]
    inter = iw * ih
    if criterion == -1:
        ua = ba + qa - inter
    elif criterion == 0:
        ua = np.broadcast_to(ba, inter.shape)
    elif criterion == 1:
        ua = np.broadcast_to(qa, inter.shape)
    else:
        ua = np.ones_like(inter)
    with np.errstate(divide='ignore', invalid='ignore'):
        out = np.where((iw > 0) & (ih > 0), inter / ua, 0.0)
    return out.astype(b.dtype)
