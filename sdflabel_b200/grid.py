"""Grid3D on the device.

Mirror of the reference's sdfrenderer/grid.py:17-71: same constructor, ``.points``
attribute (leaf tensor with a gradient hook) and ``get_surface_points`` contract.
The lattice comes from ``sdfr_lattice_points`` (bit-identical to the numpy
construction), the zero-isosurface projection / band select / order-preserving
compaction from ``sdfr_surface_extract``.
"""
from __future__ import annotations

import torch

from . import _lib

# Store grads for normals (module-global like the reference, grid.py:6; each
# Grid3D additionally keeps its own copy so two grids can be alive at once).
grads = {}


class _SurfaceExtract(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, sdf, raw_grad, threshold):
        lib = _lib.load()
        n = points.shape[0]
        dev = points.device
        pts32 = points.detach().contiguous().float()
        sdf32 = sdf.detach().contiguous().float().view(-1)
        g32 = raw_grad.detach().contiguous().float()
        out_pts = torch.empty((n, 3), device=dev, dtype=torch.float32)
        out_nrm = torch.empty((n, 3), device=dev, dtype=torch.float32)
        out_idx = torch.empty((n,), device=dev, dtype=torch.int32)
        count = torch.zeros((1,), device=dev, dtype=torch.int32)
        scratch = torch.empty((n // 1024 + 4,), device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            _lib.check(lib.sdfr_surface_extract(
                pts32.data_ptr(), 0, sdf32.data_ptr(), g32.data_ptr(), 3, 0, n, float(threshold),
                out_pts.data_ptr(), out_nrm.data_ptr(), out_idx.data_ptr(), count.data_ptr(), scratch.data_ptr(),
                _lib.stream_ptr()))
        m = int(count.item())        # dynamic output shape: one D2H of 4 bytes
        idx = out_idx[:m].long()
        nrm = out_nrm[:m]
        ctx.save_for_backward(idx, nrm)
        ctx.n = n
        ctx.mark_non_differentiable(nrm)
        return out_pts[:m].to(sdf.dtype), nrm.to(sdf.dtype)

    @staticmethod
    def backward(ctx, g_pts, _g_nrm):
        idx, nrm = ctx.saved_tensors
        g = g_pts.float()
        # p = x - f * n_hat, n_hat constant (grid.py:57-61)
        g_points = torch.zeros((ctx.n, 3), device=g.device, dtype=g.dtype).index_copy_(0, idx, g)
        g_sdf = torch.zeros((ctx.n, 1), device=g.device, dtype=g.dtype).index_copy_(
            0, idx, -(g * nrm).sum(1, keepdim=True))
        return g_points, g_sdf, None, None


class Grid3D:
    def __init__(self, density=30, device='cpu', precision=torch.float32):
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.SdfrError("sdflabel_b200.Grid3D needs a CUDA device (no CPU path); got %s" % device)
        self.density = int(density)
        pts = self.generate_point_grid(density, device)
        self.points = pts.to(device, precision).requires_grad_(True)
        self.points.register_hook(self._save_grad)

    def _save_grad(self, grad):
        self._grad = grad
        grads['grid_points'] = grad

    def generate_point_grid(self, grid_density, device=None):
        """(D^3, 3) fp32 lattice, z fastest, odd rows shifted by half a cell in x,y (grid.py:22-41)."""
        lib = _lib.load()
        _lib.require_cuda()
        device = torch.device(device if device is not None else 'cuda')
        d = int(grid_density)
        pts = torch.empty((d * d * d, 3), device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            _lib.check(lib.sdfr_lattice_points(d, pts.data_ptr(), _lib.stream_ptr()))
        return pts

    def get_surface_points(self, pred_sdf_grid, threshold=0.03):
        """Zero isosurface projection: returns projected points (M,3), NOCS (M,3), normals (M,3)."""
        # normals = d sdf / d points through the decoder's autograd node (grid.py:55-56)
        self._grad = None
        pred_sdf_grid.sum().backward(retain_graph=True)
        if self._grad is None:
            raise RuntimeError("pred_sdf_grid does not depend on this grid's points")
        points_masked, normals_masked = _SurfaceExtract.apply(self.points, pred_sdf_grid, self._grad, threshold)
        points_masked_normed = (points_masked + 1) / 2
        return points_masked, points_masked_normed, normals_masked
