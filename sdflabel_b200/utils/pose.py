"""Initial pose estimation (RANSAC over NOCS correspondences) with the searches on the device.

Mirror of the reference's utils/pose.py: ``PoseEstimator`` (7-233) with the same constructor,
``estimate`` / ``init_pose_2d`` / ``init_pose_3d`` signatures and pose dictionaries, plus
``solvePnP`` (236-283), ``procrustes`` (286-326) and ``kabsch`` (329-356).

What moves to the GPU is everything that scales with the clouds.  The reference builds two
sklearn KD-trees and, for each of its 567 hypotheses, transforms the scene cloud and queries
its nearest model points on the CPU, one hypothesis after the other.  Here

* ``sdfr_nn_query`` finds the NOCS correspondence of EVERY scene point once (the hypotheses
  only sample from it, and the final fit re-uses it);
* ``sdfr_ransac_score`` scores all surviving hypotheses in one launch (transform, exact 1-NN,
  metric + NOCS inlier test, per-hypothesis counts and masks).

The 4-point Kabsch / Procrustes fits between the two launches stay on the host in numpy: they
are O(1) per hypothesis and keeping LAPACK's SVD keeps the hypotheses bit-identical to the
reference's.  The random samples are drawn with the reference's own ``np.random.choice`` call
sequence, so a seeded run selects the same hypotheses and leaves the numpy RNG in the same state.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def _device():
    if not torch.cuda.is_available():
        raise _lib.SdfrError("PoseEstimator.init_pose_3d needs a CUDA device (there is no CPU path)")
    return torch.device('cuda', torch.cuda.current_device())


def _f32_dev(x, dev):
    if isinstance(x, torch.Tensor):
        return x.detach().to(dev, torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).to(dev)


def nn_query(queries: torch.Tensor, refs: torch.Tensor):
    """Exact nearest neighbour of ``queries`` (q,3) in ``refs`` (m,3) on the device:
    (dist float64 (q,), idx int32 (q,)), what ``KDTree(refs).query(queries)`` returns."""
    q, m = int(queries.shape[0]), int(refs.shape[0])
    idx = torch.empty((q,), device=queries.device, dtype=torch.int32)
    dist = torch.empty((q,), device=queries.device, dtype=torch.float64)
    with torch.cuda.device(queries.device):
        _lib.check(_lib.load().sdfr_nn_query(queries.data_ptr(), q, refs.data_ptr(), m, idx.data_ptr(),
                                             dist.data_ptr(), _lib.stream_ptr()))
    return dist, idx


def ransac_score(scene_pts, scene_cls, model_pts, model_cls, transforms, metric_thr, nocs_thr):
    """counts (h,) int32 and masks (h, s) uint8 of the hypotheses ``transforms`` (h,3,4)."""
    h, s = int(transforms.shape[0]), int(scene_pts.shape[0])
    counts = torch.empty((h,), device=scene_pts.device, dtype=torch.int32)
    masks = torch.empty((h, s), device=scene_pts.device, dtype=torch.uint8)
    with torch.cuda.device(scene_pts.device):
        _lib.check(_lib.load().sdfr_ransac_score(
            scene_pts.data_ptr(), scene_cls.data_ptr(), s, model_pts.data_ptr(), model_cls.data_ptr(),
            int(model_pts.shape[0]), transforms.data_ptr(), h, float(metric_thr), float(nocs_thr),
            counts.data_ptr(), masks.data_ptr(), _lib.stream_ptr()))
    return counts, masks


def legacy_choice4(n: int, draws: int, rng=None) -> np.ndarray:
    """``draws`` x ``np.random.choice(range(n), 4, replace=False)`` on numpy's global legacy generator (or on the
    ``np.random.RandomState`` given as ``rng``), (draws, 4).  Same numbers, same generator state afterwards; the
    stream is consumed by ``sdfr_np_choice4`` (a host loop over the MT19937 state) instead of 567 Python-level calls."""
    import ctypes as C
    rs = np.random if rng is None else rng
    state = rs.get_state(legacy=True)
    if state[0] != 'MT19937' or n < 4:
        return np.stack([rs.choice(n, 4, replace=False) for _ in range(draws)])
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = C.c_int32(int(state[2]))
    out = np.empty((draws, 4), dtype=np.int32)
    _lib.check(_lib.load().sdfr_np_choice4(key.ctypes.data, C.byref(pos), int(n), int(draws), out.ctypes.data))
    rs.set_state(('MT19937', key, pos.value, state[3], state[4]))
    return out.astype(np.int64)


class PoseEstimator:
    def __init__(self, type='kabsch', scale=2.2):
        self.scale = scale
        self.type = type

    def estimate(self, pcd_dsdf, nocs_dsdf, pcd_scene, nocs_scene, off_intrinsics, nocs_pred_resized, rng=None):
        """Pose dictionary (or None) for the configured estimator type, arguments as in the reference.  ``rng``
        (an ``np.random.RandomState``) replaces numpy's global generator as the source of the RANSAC samples, so
        that detections can be initialised by several threads."""
        if self.type in ('kabsch', 'procrustes'):
            return self.init_pose_3d(pcd_dsdf, nocs_dsdf, pcd_scene, nocs_scene, type=self.type,
                                     scale_model=self.scale, rng=rng)
        if self.type == 'pnp':
            return self.init_pose_2d(off_intrinsics, nocs_pred_resized, scale_model=self.scale)
        raise ValueError(f"unknown pose estimator type {self.type!r}")

    @staticmethod
    def init_pose_2d(cam, nocs_region, scale_model=1):
        """PnP on the NOCS image: a single OpenCV library call on the host, as in the reference (41-81)."""
        from scipy.spatial.transform import Rotation
        region = nocs_region.detach().cpu().permute(1, 2, 0).numpy()
        nonzero = region[:, :, 0] > 0
        object_points = (region[nonzero] * 2 - 1) * scale_model
        rows, cols = np.nonzero(nonzero)
        image_points = np.stack([rows, cols], 1).astype(np.float64)
        predicted = solvePnP(cam.cpu().numpy(), image_points, object_points)
        rot = predicted[:3, :3]
        quat = Rotation.from_matrix(rot).as_quat()        # scipy renamed from_dcm -> from_matrix
        return {'rot': rot, 'quat': np.concatenate([quat[3:], quat[:3]]), 'tra': predicted[:3, 3],
                'scale': scale_model}

    @staticmethod
    def init_pose_3d(model_pts, model_cls, scene_pts, scene_cls, metric_distance_threshold=0.15,
                     nocs_distance_threshold=0.15, type='procrustes', scale_model=1, rng=None):
        """Kabsch / Procrustes RANSAC (reference 84-233): {'scale', 'rot', 'tra'} or None."""
        dev = _device()
        scene_pts_d, scene_cls_d = _f32_dev(scene_pts, dev), _f32_dev(scene_cls, dev)
        model_pts_d, model_cls_d = _f32_dev(model_pts, dev), _f32_dev(model_cls, dev)
        total = int(scene_pts_d.shape[0])
        if total < 5:
            return None
        if type == 'kabsch':
            # the reference scales its numpy view in place; the caller's tensor is left alone here
            model_pts_d = model_pts_d * float(scale_model)

        iters = int(round((np.log(1.0 - 0.99) / np.log(1 - pow(1 - 0.7, 4))) + 0.5))     # 567
        min_num_inliers = 5

        # NOCS correspondence of every scene point, once (the reference queries 4 per hypothesis)
        cdist_d, cidx_d = nn_query(scene_cls_d, model_cls_d)
        # host copies: the scene cloud usually IS a host array; distance, index and the model cloud ride on one
        # synchronisation (the copies are enqueued in order, the last .cpu() waits for all of them)
        scene_pts_h = scene_pts_d.cpu().numpy() if isinstance(scene_pts, torch.Tensor) else \
            np.ascontiguousarray(np.asarray(scene_pts, dtype=np.float32))
        cdist_t = cdist_d.to('cpu', non_blocking=True)
        cidx_t = cidx_d.to('cpu', non_blocking=True)
        model_pts_h = model_pts_d.cpu().numpy()
        torch.cuda.current_stream(dev).synchronize()
        cdist, cidx = cdist_t.numpy(), cidx_t.numpy().astype(np.int64)

        # the reference's sample sequence (one np.random.choice per iteration, whatever happens next)
        samples = legacy_choice4(total, iters, rng)
        compatible = ~(cdist[samples] > nocs_distance_threshold).any(axis=1)

        cand = np.nonzero(compatible)[0]
        transforms = None
        if type == 'kabsch' and len(cand):
            # all 4-point fits in one batched LAPACK call (same gufunc as the per-sample calls: same bits)
            rots, tras = kabsch_batch(scene_pts_h[samples[cand]], model_pts_h[cidx[samples[cand]]])
            transforms = np.zeros((len(cand), 3, 4), dtype=np.float32)
            transforms[:, :, :3] = rots
            transforms[:, :, 3] = tras
        else:
            rows = []
            for it in cand:
                result = procrustes(scene_pts_h[samples[it]], model_pts_h[cidx[samples[it]]])
                if result is None:
                    continue
                scale, rot, tra = result
                if scale > 3:
                    continue
                trans = np.zeros((3, 4), dtype=np.float32)
                trans[:3, :3] = rot * scale
                trans[:3, 3] = tra
                rows.append(trans)
            if rows:
                transforms = np.stack(rows)
        if transforms is None or not len(transforms):
            return None

        t_d = torch.from_numpy(transforms).to(dev)
        counts_d, masks_d = ransac_score(scene_pts_d, scene_cls_d, model_pts_d, model_cls_d, t_d,
                                         metric_distance_threshold, np.float32(nocs_distance_threshold))
        counts = counts_d.cpu().numpy()
        best = int(np.argmax(counts))            # first maximum = the reference's strict '>' update
        if counts[best] < min_num_inliers:
            return None
        inliers = np.nonzero(masks_d[best].cpu().numpy())[0]

        sel_scene = scene_pts_h[inliers]
        sel_model = model_pts_h[cidx[inliers]]
        if type == 'procrustes':
            scale, rot, tra = procrustes(sel_model, sel_scene)
        else:
            rot, tra = kabsch(sel_model, sel_scene)
            scale = scale_model
        return {'scale': scale, 'rot': rot, 'tra': tra}


def solvePnP(cam, image_points, object_points, return_inliers=False):
    """OpenCV RANSAC PnP with the reference's settings (236-283): 1000 iterations, 1 px reprojection error."""
    import cv2
    pose, inliers = np.eye(4), []
    if image_points.shape[0] >= 4:
        image_points[:, [0, 1]] = image_points[:, [1, 0]]
        ok, rvec, tvec, inliers = cv2.solvePnPRansac(
            np.expand_dims(object_points, 1), np.expand_dims(image_points, 1).astype(float), cam, np.zeros((4, 1)),
            iterationsCount=1000, reprojectionError=1.)[:4]
        if ok:
            pose[:3, :3] = cv2.Rodrigues(rvec)[0]
            pose[:3, 3] = np.squeeze(tvec)
        if inliers is None:
            inliers = []
    return (pose, len(inliers)) if return_inliers else pose


def procrustes(from_points, to_points):
    """Similarity transform (c, R, t) with to ~ c R from + t (Umeyama); None for a degenerate sample."""
    assert from_points.ndim == 2 and from_points.shape == to_points.shape
    n, m = from_points.shape
    mu_f, mu_t = from_points.mean(axis=0), to_points.mean(axis=0)
    d_f, d_t = from_points - mu_f, to_points - mu_t
    var_f = (d_f * d_f).sum(axis=1).mean()
    cov = d_t.T.dot(d_f) / n
    u, d, vt = np.linalg.svd(cov, full_matrices=True)
    rank = np.linalg.matrix_rank(cov)
    s = np.eye(m)
    if rank >= m - 1 and np.linalg.det(cov) < 0:
        s[m - 1, m - 1] = -1
    elif rank < m - 1:
        return None
    r = u.dot(s).dot(vt)
    c = (d * s.diagonal()).sum() / var_f
    return c, r, mu_t - c * r.dot(mu_f)


def kabsch_batch(canonical_points, predicted_points):
    """``kabsch`` for stacks (h,4,3) of samples: the reductions, products and the SVD are numpy's batched
    forms of the very same calls."""
    mu_c, mu_p = np.mean(canonical_points, axis=1), np.mean(predicted_points, axis=1)
    c_c = canonical_points - mu_c[:, None, :]
    p_c = predicted_points - mu_p[:, None, :]
    u, _, vt = np.linalg.svd(np.matmul(p_c.transpose(0, 2, 1), c_c))
    rot = np.matmul(u, vt)
    neg = np.linalg.det(rot) < 0.0
    if neg.any():
        vt[neg, -1, :] *= -1.0
        rot[neg] = np.matmul(u[neg], vt[neg])
    # the reference's translation, np.dot(R, p - c) - np.dot(R, p) + p, for the whole stack: matmul hands every
    # 3x3 . 3 product to the same BLAS gemv as np.dot does (bit-identical, tests/test_pose_oracle.py)
    t = mu_p - mu_c
    tras = np.matmul(rot, t[:, :, None])[:, :, 0] - np.matmul(rot, mu_p[:, :, None])[:, :, 0] + mu_p
    return rot, tras


def kabsch(canonical_points, predicted_points):
    """Rigid fit (R, t) of the reference's Kabsch step, including its translation formula."""
    mu_c, mu_p = np.mean(canonical_points, axis=0), np.mean(predicted_points, axis=0)
    c_c = canonical_points - np.expand_dims(mu_c, axis=0)
    p_c = predicted_points - np.expand_dims(mu_p, axis=0)
    u, _, vt = np.linalg.svd(p_c.T @ c_c)
    rot = u @ vt
    if np.linalg.det(rot) < 0.0:
        vt[-1, :] *= -1.0
        rot = np.dot(u, vt)
    tra = mu_p - mu_c
    tra = np.dot(rot, tra) - np.dot(rot, mu_p) + mu_p
    return rot, tra
