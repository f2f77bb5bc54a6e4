"""CPU restatement of the reference's rotated-box overlap (TEST INFRASTRUCTURE - not shipped code).

Follows pipelines/rotate_iou.py line by line in numpy float32 scalar arithmetic: ``rbbox_to_corners``
(203-226), ``point_in_quadrilateral`` (158-175), ``line_segment_intersection`` (76-116),
``quadrilateral_intersection`` (178-200), ``sort_vertex_in_convex_polygon`` (35-72), ``area`` (27-31),
``devRotateIoUEval`` (245-254) and the host helpers ``d3_box_overlap_kernel`` (328-357) and
``image_box_overlap`` (360-379).  Pinned against the unmodified reference kernel run under numba's CUDA
simulator: ``oracle/make_golden_iou.py`` -> ``tests/golden/rotate_iou.npz``.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def corners(b):
    c, s = math.cos(b[4]), math.sin(b[4])          # the simulator evaluates these in double
    xs = [-b[2] / 2, -b[2] / 2, b[2] / 2, b[2] / 2]
    ys = [-b[3] / 2, b[3] / 2, b[3] / 2, -b[3] / 2]
    out = np.zeros(8, F)
    for i in range(4):
        out[2 * i] = c * xs[i] + s * ys[i] + b[0]
        out[2 * i + 1] = -s * xs[i] + c * ys[i] + b[1]
    return out


def point_in_quad(px, py, c):
    ab0, ab1 = c[2] - c[0], c[3] - c[1]
    ad0, ad1 = c[6] - c[0], c[7] - c[1]
    ap0, ap1 = px - c[0], py - c[1]
    abab, abap = ab0 * ab0 + ab1 * ab1, ab0 * ap0 + ab1 * ap1
    adad, adap = ad0 * ad0 + ad1 * ad1, ad0 * ap0 + ad1 * ap1
    eps = 0.0001
    return abab >= abap - eps and abap >= 0 - eps and adad >= adap - eps and adap >= 0 - eps


def segment_intersection(p1, p2, i, j):
    a = (p1[2 * i], p1[2 * i + 1])
    b = (p1[2 * ((i + 1) % 4)], p1[2 * ((i + 1) % 4) + 1])
    c = (p2[2 * j], p2[2 * j + 1])
    d = (p2[2 * ((j + 1) % 4)], p2[2 * ((j + 1) % 4) + 1])
    ba0, ba1 = b[0] - a[0], b[1] - a[1]
    da0, ca0, da1, ca1 = d[0] - a[0], c[0] - a[0], d[1] - a[1], c[1] - a[1]
    acd = da1 * ca0 > ca1 * da0
    bcd = (d[1] - b[1]) * (c[0] - b[0]) > (c[1] - b[1]) * (d[0] - b[0])
    if acd != bcd:
        abc = ca1 * ba0 > ba1 * ca0
        abd = da1 * ba0 > ba1 * da0
        if abc != abd:
            dc0, dc1 = d[0] - c[0], d[1] - c[1]
            abba = a[0] * b[1] - b[0] * a[1]
            cddc = c[0] * d[1] - d[0] * c[1]
            dh = ba1 * dc0 - ba0 * dc1
            return F((abba * dc0 - ba0 * cddc) / dh), F((abba * dc1 - ba1 * cddc) / dh)
    return None


def sort_vertices(pts, n):
    if n <= 0:
        return
    cx = cy = F(0)
    for i in range(n):
        cx, cy = F(cx + pts[2 * i]), F(cy + pts[2 * i + 1])
    cx, cy = F(cx / n), F(cy / n)
    vs = np.zeros(16, F)
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in range(n):
            vx, vy = F(pts[2 * i] - cx), F(pts[2 * i + 1] - cy)
            d = F(math.sqrt(vx * vx + vy * vy))
            vx, vy = F(vx / d), F(vy / d)
            if vy < 0:
                vx = F(-2 - vx)
            vs[i] = vx
    for i in range(1, n):
        if vs[i - 1] > vs[i]:
            temp, tx, ty = vs[i], pts[2 * i], pts[2 * i + 1]
            j = i
            while j > 0 and vs[j - 1] > temp:
                vs[j] = vs[j - 1]
                pts[2 * j], pts[2 * j + 1] = pts[2 * j - 2], pts[2 * j - 1]
                j -= 1
            vs[j] = temp
            pts[2 * j], pts[2 * j + 1] = tx, ty


def triangle_area(a, b, c):
    return ((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0


def intersection_area(b1, b2):
    c1, c2 = corners(b1), corners(b2)
    pts = np.zeros(16, F)
    n = 0
    for i in range(4):
        if point_in_quad(c1[2 * i], c1[2 * i + 1], c2):
            pts[2 * n], pts[2 * n + 1] = c1[2 * i], c1[2 * i + 1]
            n += 1
        if point_in_quad(c2[2 * i], c2[2 * i + 1], c1):
            pts[2 * n], pts[2 * n + 1] = c2[2 * i], c2[2 * i + 1]
            n += 1
    for i in range(4):
        for j in range(4):
            t = segment_intersection(c1, c2, i, j)
            if t is not None:
                pts[2 * n], pts[2 * n + 1] = t       # (the reference has no bound check either)
                n += 1
    sort_vertices(pts, n)
    area = 0.0
    for i in range(n - 2):
        area += abs(triangle_area(pts[:2], pts[2 * i + 2:2 * i + 4], pts[2 * i + 4:2 * i + 6]))
    return area


def rotate_iou(boxes, query, criterion=-1):
    boxes, query = np.asarray(boxes, F), np.asarray(query, F)
    out = np.zeros((boxes.shape[0], query.shape[0]), F)
    for n in range(boxes.shape[0]):
        for k in range(query.shape[0]):
            r1, r2 = query[k], boxes[n]                 # argument order of rotate_iou.py:286
            a1, a2 = r1[2] * r1[3], r2[2] * r2[3]
            inter = intersection_area(r1, r2)
            if criterion == -1:
                v = inter / (a1 + a2 - inter)
            elif criterion == 0:
                v = inter / a1
            elif criterion == 1:
                v = inter / a2
            else:
                v = inter
            out[n, k] = v
    return out


def d3_box_overlap(boxes, qboxes, rinc, criterion=-1, camera_coordinate=False):
    """rotate_iou.py:328-357 (returns the updated copy of ``rinc``)."""
    rinc = np.array(rinc, copy=True)
    for i in range(boxes.shape[0]):
        for j in range(qboxes.shape[0]):
            if rinc[i, j] > 0:
                if camera_coordinate:
                    iw = min(boxes[i, 1], qboxes[j, 1]) - max(boxes[i, 1] - boxes[i, 4], qboxes[j, 1] - qboxes[j, 4])
                else:
                    iw = min(boxes[i, 2] + boxes[i, 5], qboxes[j, 2] + qboxes[j, 5]) - max(boxes[i, 2], qboxes[j, 2])
                if iw > 0:
                    area1 = boxes[i, 3] * boxes[i, 4] * boxes[i, 5]
                    area2 = qboxes[j, 3] * qboxes[j, 4] * qboxes[j, 5]
                    inc = iw * rinc[i, j]
                    ua = (area1 + area2 - inc) if criterion == -1 else area1 if criterion == 0 else area2 if criterion == 1 else inc
                    rinc[i, j] = inc / ua
                else:
                    rinc[i, j] = 0.0
    return rinc


def image_box_overlap(boxes, query_boxes, criterion=-1):
    """rotate_iou.py:360-379."""
    N, K = boxes.shape[0], query_boxes.shape[0]
    overlaps = np.zeros((N, K), dtype=boxes.dtype)
    for k in range(K):
        qa = (query_boxes[k, 2] - query_boxes[k, 0]) * (query_boxes[k, 3] - query_boxes[k, 1])
        for n in range(N):
            iw = min(boxes[n, 2], query_boxes[k, 2]) - max(boxes[n, 0], query_boxes[k, 0])
            if iw > 0:
                ih = min(boxes[n, 3], query_boxes[k, 3]) - max(boxes[n, 1], query_boxes[k, 1])
                if ih > 0:
                    ba = (boxes[n, 2] - boxes[n, 0]) * (boxes[n, 3] - boxes[n, 1])
                    ua = (ba + qa - iw * ih) if criterion == -1 else ba if criterion == 0 else qa if criterion == 1 else 1.0
                    overlaps[n, k] = iw * ih / ua
    return overlaps
