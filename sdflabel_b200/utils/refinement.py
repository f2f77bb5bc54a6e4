"""The two helpers of the reference's utils/refinement.py that touch the refine path.

``rot_from_yaw`` (reference utils/refinement.py:108-125) is folded into the pose kernel of the
fused loop; this host-side copy exists for callers that build a render pose themselves.
``get_kitti_label`` (501-562) is the dump-time label reconstruction: one more lattice
evaluation + surface extraction on the device, the rest is 4x4 bookkeeping on the host.
Everything else in that file (IoU, frustum, open3d line sets, LIDAR depth maps) is host-side
KITTI bookkeeping outside the hot path.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def rot_from_yaw(yaw):
    """3x3 rotation about the y axis: [[c,0,s],[0,1,0],[-s,0,c]]."""
    if not isinstance(yaw, torch.Tensor):
        yaw = torch.tensor([float(yaw)])
    c, s = torch.cos(yaw).reshape(()), torch.sin(yaw).reshape(())
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([c, z, s, z, o, z, -s, z, c]).view(3, 3)


def roty_in_bev(pose):
    """KITTI rotation_y (x right, y down, z forward; BEV = x-z plane) of a 4x4 / 3x3 object pose:
    the angle between the x axis and the object's rotated z axis, negative when that axis points
    forward (reference utils/refinement.py:201-220)."""
    fwd = np.asarray(pose)[:3, 2]                       # R @ [0, 0, 1]
    rot_y = math.acos(float(np.clip(fwd[0], -1.0, 1.0)))
    return -rot_y if fwd[2] > 0 else rot_y


def alpha_in_bev(pose, rot_y):
    """KITTI observation angle: rotation_y corrected by the bearing of the object's position in the
    BEV plane (reference utils/refinement.py:223-257)."""
    x, z = float(pose[0, 3]), float(pose[2, 3])
    theta = math.atan2(abs(x), abs(z))
    return rot_y + theta if x < 0 else rot_y - theta


def get_kitti_label(dsdf, grid, latent, scale, trans, yaw, p_WC, bbox):
    """KITTI label from the refined parameters (same arguments as the reference).  The extent
    comes from the scaled zero-isosurface points of the final, un-normalised latent
    (refine_css.py:229 -> refinement.py:536), evaluated by the same kernels as the loop."""
    precision, device = grid.points.dtype, grid.points.device
    results = {'yaw': yaw.detach().cpu().numpy(), 'trans': trans.detach().cpu().numpy(),
               'scale': scale.detach().cpu().numpy(), 'latent': latent.detach().cpu().numpy()}
    cam_T = np.eye(4)
    cam_T[:3, :3] = rot_from_yaw(float(results['yaw'].reshape(-1)[0])).numpy() @ np.diag([1, -1, 1])
    cam_T[:3, 3] = results['trans'] * results['scale']
    global_T = np.linalg.inv(p_WC) @ cam_T

    inputs = torch.cat([latent.to(device, precision).expand(grid.points.size(0), -1), grid.points], 1)
    pred_sdf_grid, _ = dsdf(inputs)
    points_masked, _, _ = grid.get_surface_points(pred_sdf_grid)
    scaled_points = points_masked.detach().cpu().numpy() * results['scale'][None]
    mins, maxs = scaled_points.min(0), scaled_points.max(0)
    width, height, length = (maxs - mins).tolist()
    bottom_center = np.asarray([0, mins[1], 0])

    label = {'name': 'Car', 'bbox': bbox}
    label['location'] = (global_T[:3, :3] @ bottom_center.T).T + global_T[:3, 3]
    label['dimensions'] = [height, width, length]
    label['rotation_y'] = roty_in_bev(global_T)
    label['alpha'] = alpha_in_bev(global_T, label['rotation_y'])
    label['score'] = 1
    return label, scaled_points, cam_T


def kitti_label_from_extents(ext_min, ext_max, latent, scale, trans, yaw, p_WC, bbox):
    """``get_kitti_label`` with the extents already known: ``ext_min`` / ``ext_max`` are the bounds of the
    un-scaled isosurface points of the raw latent (``sdfr_refine_label_extents``, evaluated for all detections of
    a batch in one pass).  min / max commute with the multiplication by the positive float32 scale, so
    ``ext * scale`` is bit for bit what the reference gets from ``(points * scale).min()`` (refinement.py:536-541).
    Returns (label, cam_T) - the scaled point cloud itself only feeds the reference's visualiser."""
    scale32 = np.asarray(scale, dtype=np.float32).reshape(-1)[:1]
    yaw_f = float(np.asarray(yaw, dtype=np.float32).reshape(-1)[0])
    trans32 = np.asarray(trans, dtype=np.float32).reshape(3)
    cam_T = np.eye(4)
    cam_T[:3, :3] = rot_from_yaw(yaw_f).numpy() @ np.diag([1, -1, 1])
    cam_T[:3, 3] = trans32 * scale32
    global_T = np.linalg.inv(p_WC) @ cam_T
    mins = np.asarray(ext_min, dtype=np.float32) * scale32
    maxs = np.asarray(ext_max, dtype=np.float32) * scale32
    width, height, length = (maxs - mins).tolist()
    bottom_center = np.asarray([0, mins[1], 0])
    label = {'name': 'Car', 'bbox': bbox}
    label['location'] = (global_T[:3, :3] @ bottom_center.T).T + global_T[:3, 3]
    label['dimensions'] = [height, width, length]
    label['rotation_y'] = roty_in_bev(global_T)
    label['alpha'] = alpha_in_bev(global_T, label['rotation_y'])
    label['score'] = 1
    return label, cam_T
