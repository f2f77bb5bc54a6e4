"""Host-side helpers of the rasteriser.

Mirror of ``calibration_matrix`` (reference sdfrenderer/renderer/utils_rasterer.py:59-83)
and ``qrot`` (6-24).  ``qrot`` here is a plain torch expression kept for callers
that build poses on the host; the kernels carry their own copy (csrc/splat.cu).
The HPR helpers (sphericalFlip / convexHull, 27-56) are never enabled by any
caller of the refine path and are out of scope.
"""
import numpy as np
import torch


def qrot(q, v):
    """Rotate v (*, 3) by quaternion q (*, 4) = [w, x, y, z]; q is not normalised."""
    if q.shape[-1] != 4 or v.shape[-1] != 3 or q.shape[:-1] != v.shape[:-1]:
        raise ValueError("qrot expects q (*,4) and v (*,3) with matching leading dims")
    u = q[..., 1:]
    t = torch.linalg.cross(u, v, dim=-1)
    return v + 2 * (q[..., :1] * t + torch.linalg.cross(u, t, dim=-1))


def calibration_matrix(resolution_px, diagonal_mm, focal_len_mm, skew=0.):
    """K from sensor geometry: pixels-per-mm follow from the sensor diagonal, the
    principal point is the image centre."""
    w_px, h_px = resolution_px
    px_per_mm = float(np.hypot(w_px, h_px)) / diagonal_mm
    f_px = focal_len_mm * px_per_mm
    return np.array([[f_px, skew, w_px / 2], [0, f_px, h_px / 2], [0, 0, 1]])
