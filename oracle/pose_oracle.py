"""CPU restatement of the reference's initial-pose RANSAC (TEST INFRASTRUCTURE - not shipped code).

Follows ``PoseEstimator.init_pose_3d`` (reference utils/pose.py:84-233), ``procrustes``
(286-326) and ``kabsch`` (329-356) line by line, with the two sklearn KD-trees replaced by an
exact brute-force nearest-neighbour search (float64 distance on the float32 coordinates, which
is what a KD-tree returns).  The random samples are drawn with the same ``np.random.choice``
calls in the same order, so a seeded run consumes the global numpy RNG exactly like the
reference does.  Pinned against the unmodified reference by ``oracle/make_golden_pose.py`` ->
``tests/golden/pose_ransac.npz``.

Besides the pose it returns the per-hypothesis record (valid flag, transform, inlier count) the
CUDA path is checked against.
"""
from __future__ import annotations

import numpy as np


def nn_exact(queries: np.ndarray, refs: np.ndarray):
    """(dist float64 [q], idx [q]) of the exact nearest neighbour; ties -> lowest index."""
    q = np.asarray(queries, dtype=np.float32).astype(np.float64)
    r = np.asarray(refs, dtype=np.float32).astype(np.float64)
    idx = np.empty(q.shape[0], dtype=np.int64)
    dist = np.empty(q.shape[0], dtype=np.float64)
    for s in range(0, q.shape[0], 512):
        d = q[s:s + 512, None, :] - r[None, :, :]
        d2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]   # KD-tree order, no FMA
        i = d2.argmin(1)
        idx[s:s + 512] = i
        dist[s:s + 512] = np.sqrt(d2[np.arange(d2.shape[0]), i])
    return dist, idx


def kabsch(canonical_points, predicted_points):
    """pose.py:329-356."""
    canonical_mean = np.mean(canonical_points, axis=0)
    predicted_mean = np.mean(predicted_points, axis=0)
    canonical_centered = canonical_points - np.expand_dims(canonical_mean, axis=0)
    predicted_centered = predicted_points - np.expand_dims(predicted_mean, axis=0)
    cross_correlation = predicted_centered.T @ canonical_centered
    u, s, vt = np.linalg.svd(cross_correlation)
    rotation = u @ vt
    if np.linalg.det(rotation) < 0.0:
        vt[-1, :] *= -1.0
        rotation = np.dot(u, vt)
    translation = predicted_mean - canonical_mean
    translation = np.dot(rotation, translation) - np.dot(rotation, predicted_mean) + predicted_mean
    return rotation, translation


def procrustes(from_points, to_points):
    """pose.py:286-326."""
    N, m = from_points.shape
    mean_from = from_points.mean(axis=0)
    mean_to = to_points.mean(axis=0)
    delta_from = from_points - mean_from
    delta_to = to_points - mean_to
    sigma_from = (delta_from * delta_from).sum(axis=1).mean()
    cov_matrix = delta_to.T.dot(delta_from) / N
    U, d, V_t = np.linalg.svd(cov_matrix, full_matrices=True)
    cov_rank = np.linalg.matrix_rank(cov_matrix)
    S = np.eye(m)
    if cov_rank >= m - 1 and np.linalg.det(cov_matrix) < 0:
        S[m - 1, m - 1] = -1
    elif cov_rank < m - 1:
        return None
    R = U.dot(S).dot(V_t)
    c = (d * S.diagonal()).sum() / sigma_from
    t = mean_to - c * R.dot(mean_from)
    return c, R, t


def num_ransac_iterations(p=0.99, outlier_prob=0.7, sample_size=4) -> int:
    """pose.py:122-128 (= 567)."""
    return int(round((np.log(1.0 - p) / np.log(1 - pow(1 - outlier_prob, sample_size))) + 0.5))


def init_pose_3d(model_pts, model_cls, scene_pts, scene_cls, metric_distance_threshold=0.15,
                 nocs_distance_threshold=0.15, type='procrustes', scale_model=1, return_trace=False, nn='brute'):
    """``nn='kdtree'`` uses sklearn KD-trees like the reference (the timed CPU baseline of
    tools/pose_bench.py); ``'brute'`` is the dependency-free exact search the parity tests use."""
    model_pts = np.array(model_pts, copy=True)
    model_cls = np.asarray(model_cls)
    scene_pts = np.asarray(scene_pts)
    scene_cls = np.asarray(scene_cls)
    if scene_pts.shape[0] < 5:                                                  # pose.py:116
        return (None, None) if return_trace else None
    if type == 'kabsch':
        model_pts *= scale_model                                                # pose.py:119-120
    total = scene_pts.shape[0]
    iters = num_ransac_iterations()
    if nn == 'kdtree':
        from sklearn.neighbors import KDTree
        tree_c, tree_p = KDTree(model_cls), KDTree(model_pts)                   # pose.py:133-134

        def _query(tree):
            def q(x, _refs):
                d, i = tree.query(x)
                return d[:, 0], i[:, 0]
            return q
        nn_cls, nn_pts = _query(tree_c), _query(tree_p)
    else:
        nn_cls = nn_pts = nn_exact
    best = np.array([], dtype=np.int64)
    trace = {'valid': np.zeros(iters, bool), 'transforms': np.zeros((iters, 12), np.float32),
             'counts': np.zeros(iters, np.int64), 'samples': np.zeros((iters, 4), np.int64)}
    for it in range(iters):
        indices = np.random.choice(range(total), 4, replace=False)             # pose.py:139
        trace['samples'][it] = indices
        sel_pts, sel_cls = scene_pts[indices], scene_cls[indices]
        dists, idxs_nocs = nn_cls(sel_cls, model_cls)                         # pose.py:146
        if (dists > nocs_distance_threshold).any():                             # pose.py:151
            continue
        sel_model = model_pts[idxs_nocs]
        if type == 'procrustes':
            result = procrustes(sel_pts, sel_model)
            if result is None:
                continue
            scale, rot, tra = result
        else:
            rot, tra = kabsch(sel_pts, sel_model)
            scale = 1
        if scale > 3:                                                           # pose.py:163
            continue
        trans = np.zeros((3, 4), dtype=np.float32)
        trans[:3, :3] = rot * scale
        trans[:3, 3] = tra
        transformed = (trans[:, :3] @ scene_pts.T).T + trans[:, 3]              # pose.py:170
        dists, idxs = nn_pts(transformed, model_pts)                          # pose.py:172
        dists_color = np.linalg.norm(scene_cls - model_cls[idxs], axis=1)
        inliers = np.where((dists < metric_distance_threshold) & (dists_color < nocs_distance_threshold))[0]
        trace['valid'][it] = True
        trace['transforms'][it] = trans.reshape(-1)
        trace['counts'][it] = len(inliers)
        if len(inliers) > len(best):                                            # pose.py:192
            best = inliers
    trace['best'] = best
    if len(best) < 5:                                                           # pose.py:196
        return (None, trace) if return_trace else None
    sel_pts, sel_cls = scene_pts[best], scene_cls[best]
    _, idxs = nn_cls(sel_cls, model_cls)                                        # pose.py:202
    sel_model = model_pts[idxs]
    if type == 'procrustes':
        scale, rot, tra = procrustes(sel_model, sel_pts)
    else:
        rot, tra = kabsch(sel_model, sel_pts)
        scale = scale_model
    pose = {'scale': scale, 'rot': rot, 'tra': tra}
    return (pose, trace) if return_trace else pose


def make_pose_scene(seed=3, n_model=1500, n_scene=300, n_outliers=60, scale=2.0, yaw=0.7,
                    tra=(1.0, 0.5, 8.0), noise=0.01):
    """Seeded synthetic input: a car-sized ellipsoid model cloud with NOCS colours and a noisy, partially
    outlying scene cloud seen under a known similarity transform (model -> scene)."""
    rng = np.random.RandomState(seed)
    u = rng.normal(size=(n_model, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    model_pts = (u * np.array([0.45, 0.3, 0.9])).astype(np.float32)
    model_cls = ((model_pts + 1) / 2).astype(np.float32)
    pick = rng.choice(n_model, n_scene, replace=False)
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    pts = (R @ (model_pts[pick].astype(np.float64) * scale).T).T + np.asarray(tra)
    pts += rng.normal(scale=noise, size=pts.shape)
    cls = model_cls[pick] + rng.normal(scale=noise, size=(n_scene, 3))
    out_pts = rng.uniform(-1.5, 1.5, size=(n_outliers, 3)) + np.asarray(tra)
    out_cls = rng.uniform(0, 1, size=(n_outliers, 3))
    scene_pts = np.concatenate([pts, out_pts]).astype(np.float32)
    scene_cls = np.concatenate([cls, out_cls]).astype(np.float32)
    perm = rng.permutation(scene_pts.shape[0])
    return {'model_pts': model_pts, 'model_cls': model_cls, 'scene_pts': scene_pts[perm],
            'scene_cls': scene_cls[perm], 'scale': scale, 'rot': R, 'tra': np.asarray(tra)}
