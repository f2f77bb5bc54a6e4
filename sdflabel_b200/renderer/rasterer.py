"""Differentiable surfel rasteriser on the device.

Mirror of ``Rasterer`` in the reference's sdfrenderer/renderer/rasterer.py:9-155
(same constructor and ``forward`` signature, same ``rendering`` / ``points``
dictionaries).  Projection (renderer/projection.py), the tangent-disc primitive
(renderer/primitives.py:165-243) and the composition are one pair of CUDA
launches behind ``sdfr_splat_forward``; the backward of the whole thing is
``sdfr_splat_backward``.  ``primitives='disc'`` without background is what the
refine path uses (optimizer.py:110-123); the screen-space circle primitives
(``'circle'``, ``'circle_opt'``, primitives.py:4-162) and background compositing
(``bg``, rasterer.py:107-111) serve the stand-alone render API of
sdfrenderer/main.py:62-121 and run through the same two entry points.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .utils_rasterer import calibration_matrix


class _Splat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, normals, colors, camera, cfg_tuple):
        lib = _lib.load()
        width, height, kinv, kmat, rot, output_nocs, primitive, bg = cfg_tuple
        dev = coords.device
        bg32 = None if bg is None else bg.detach().to(dev, torch.float32).contiguous()
        cfg = _lib.RasterCfg(width=width, height=height, rot=rot, output_nocs=int(output_nocs), primitive=primitive,
                             bg_dev=_lib.ptr(bg32))
        cfg.kinv[:] = kinv
        cfg.k[:] = kmat
        c32 = coords.detach().contiguous().float()
        n32 = normals.detach().contiguous().float()
        col32 = colors.detach().contiguous().float() if colors is not None else None
        if rot == _lib.ROT_DCM:
            pose32 = camera.detach().to(dev).float().contiguous().view(-1)
        else:
            pose32 = camera.detach().to(dev).float().contiguous().view(-1)
        m = c32.shape[0]
        P = width * height
        f32 = dict(device=dev, dtype=torch.float32)
        color = torch.empty((3, height, width), **f32)
        mask = torch.empty((1, height, width), **f32)
        # with a background only colour and mask can be composed (rasterer.py:107-144)
        depth = torch.empty((1, height, width), **f32) if bg32 is None else None
        nmap = torch.empty((3, height, width), **f32) if bg32 is None else None
        cam_pts = torch.empty((m, 3), **f32)
        cam_rgb = torch.empty((m, 3), **f32)
        front = torch.empty((m,), device=dev, dtype=torch.uint8)
        want_front = rot == _lib.ROT_DCM
        xyzf = torch.empty((m, 3), **f32) if want_front else None
        rgbf = torch.empty((m, 3), **f32) if want_front else None
        count = torch.zeros((1,), device=dev, dtype=torch.int32) if want_front else None
        ws_bytes = lib.sdfr_splat_workspace_bytes(cfg, m)
        ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(lib.sdfr_splat_forward(
                cfg, c32.data_ptr(), n32.data_ptr(), _lib.ptr(col32), pose32.data_ptr(), m,
                color.data_ptr(), mask.data_ptr(), _lib.ptr(depth), _lib.ptr(nmap), cam_pts.data_ptr(),
                cam_rgb.data_ptr(), front.data_ptr(), _lib.ptr(xyzf), _lib.ptr(rgbf), _lib.ptr(count),
                ws.data_ptr(), _lib.stream_ptr()))
        if want_front:
            mf = int(count.item())   # dynamic shape of points['xyzf']
            xyzf, rgbf = xyzf[:mf], rgbf[:mf]
            front_idx = torch.nonzero(front, as_tuple=False).view(-1)
        else:
            xyzf = torch.empty((0, 3), **f32)
            rgbf = torch.empty((0, 3), **f32)
            front_idx = torch.empty((0,), device=dev, dtype=torch.long)
        if bg32 is not None:
            depth = torch.zeros((1, height, width), **f32)
            nmap = torch.zeros((3, height, width), **f32)
        ctx.cfg = cfg
        ctx.bg32 = bg32           # keeps the buffer cfg.bg_dev points at alive until the backward
        ctx.cfg_tuple = cfg_tuple
        ctx.ws = ws
        ctx.save_for_backward(c32, n32, col32 if col32 is not None else torch.empty(0, device=dev), pose32, front_idx)
        ctx.has_colors = col32 is not None
        ctx.dtypes = (coords.dtype, normals.dtype, colors.dtype if colors is not None else None, camera.dtype)
        ctx.camera_shape = tuple(camera.shape)
        ctx.camera_device = camera.device
        dt = coords.dtype
        return (color.to(dt), mask.to(dt), depth.to(dt), nmap.to(dt), cam_pts.to(dt), cam_rgb.to(dt), xyzf.to(dt),
                rgbf.to(dt))

    @staticmethod
    def backward(ctx, g_color, g_mask, g_depth, g_nmap, g_pts, g_rgb, g_xyzf, g_rgbf):
        lib = _lib.load()
        c32, n32, col32, pose32, front_idx = ctx.saved_tensors
        width, height, kinv, kmat, rot, output_nocs, primitive, bg = ctx.cfg_tuple
        dev = c32.device
        m = c32.shape[0]

        def prep(g):
            return None if g is None else g.contiguous().float()

        g_color, g_mask, g_depth, g_nmap = prep(g_color), prep(g_mask), prep(g_depth), prep(g_nmap)
        if ctx.bg32 is not None:
            g_depth = g_nmap = None
        g_pts, g_rgb = prep(g_pts), prep(g_rgb)
        if g_xyzf is not None and front_idx.numel():
            g_pts = (g_pts if g_pts is not None else torch.zeros((m, 3), device=dev)).index_add(
                0, front_idx, g_xyzf.float())
        if g_rgbf is not None and front_idx.numel():
            g_rgb = (g_rgb if g_rgb is not None else torch.zeros((m, 3), device=dev)).index_add(
                0, front_idx, g_rgbf.float())
        d_coords = torch.empty((m, 3), device=dev, dtype=torch.float32)
        d_normals = torch.empty((m, 3), device=dev, dtype=torch.float32)
        d_colors = torch.empty((m, 3), device=dev, dtype=torch.float32) if (ctx.has_colors and not output_nocs) else None
        npose = 12 if rot == _lib.ROT_DCM else 7
        d_pose = torch.zeros((npose,), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.sdfr_splat_backward(
                ctx.cfg, c32.data_ptr(), n32.data_ptr(), col32.data_ptr() if ctx.has_colors else 0,
                pose32.data_ptr(), m, _lib.ptr(g_color), _lib.ptr(g_mask), _lib.ptr(g_depth), _lib.ptr(g_nmap),
                _lib.ptr(g_pts), _lib.ptr(g_rgb), d_coords.data_ptr(), d_normals.data_ptr(), _lib.ptr(d_colors),
                d_pose.data_ptr(), ctx.ws.data_ptr(), _lib.stream_ptr()))
        dt_c, dt_n, dt_col, dt_cam = ctx.dtypes
        if rot == _lib.ROT_DCM:
            g_cam = torch.zeros(ctx.camera_shape, device=dev, dtype=torch.float32)
            g_cam[:3, :4] = d_pose.view(3, 4)
        else:
            g_cam = d_pose
        g_cam = g_cam.to(ctx.camera_device, dt_cam)
        g_colors = None
        if ctx.has_colors:
            g_colors = d_colors.to(dt_col) if d_colors is not None else torch.zeros_like(col32).to(dt_col)
        return d_coords.to(dt_c), d_normals.to(dt_n), g_colors, g_cam, None


class Rasterer(torch.nn.Module):
    def __init__(self, K, resolution_px, diagonal_mm=20, focal_len_mm=70, precision=torch.float32):
        """
        Args:
            K (torch.Tensor): intrinsic camera parameters (3,3) or None
            resolution_px (tuple): camera resolution in pixels (W, H)
            diagonal_mm, focal_len_mm: used to build K when it is None
            precision: dtype of the stored K
        """
        super().__init__()
        self.res_x_px, self.res_y_px = resolution_px
        # same buffers as the reference keeps (rasterer.py:25-32); the kernels generate pixel
        # coordinates from the index, these only exist for API compatibility
        yy, xx = np.mgrid[0:self.res_y_px, 0:self.res_x_px]
        grid = np.concatenate((xx[..., None], yy[..., None]), axis=-1)
        self.register_buffer('grid', torch.from_numpy(grid.reshape((1, -1, 2))))
        yy, xx = np.mgrid[-7:8, -7:8]
        grid_prim = np.concatenate((xx[..., None], yy[..., None]), axis=-1)
        self.register_buffer('grid_prim', torch.from_numpy(grid_prim.reshape((1, -1, 2))))
        if K is None:
            K = torch.from_numpy(calibration_matrix(
                resolution_px=(self.res_x_px, self.res_y_px), diagonal_mm=diagonal_mm, focal_len_mm=focal_len_mm,
                skew=0))
        self.register_buffer('K', K.to(precision))
        self._k_cache = None

    def _intrinsics(self):
        """(kinv, k) as python float lists; K is inverted in fp32 on the host like the
        reference does on its device (primitives.py:204)."""
        key = (self.K.data_ptr(), self.K._version)
        if self._k_cache is None or self._k_cache[0] != key:
            k32 = self.K.detach().float().cpu()
            kinv = k32.inverse()
            self._k_cache = (key, kinv.reshape(-1).tolist(), k32.reshape(-1).tolist())
        return self._k_cache[1], self._k_cache[2]

    def forward(
        self,
        coords,
        normals,
        colors,
        camera_matrix,
        rot='quat',
        primitives='disc',
        bg=None,
        output_mask=False,
        output_depth=False,
        output_normals=False,
        output_nocs=False,
        output_points=True
    ):
        prim = {'disc': _lib.PRIM_DISC, 'circle': _lib.PRIM_CIRCLE, 'circle_opt': _lib.PRIM_CIRCLE_OPT}.get(primitives)
        if prim is None:
            raise ValueError(f"unknown primitive {primitives!r}")      # (the reference dies with UnboundLocalError)
        if bg is not None and (output_depth or output_normals):
            # rasterer.py:133-144 multiplies M+1 weight rows with M depth / normal rows
            raise RuntimeError("with a background only the color and mask maps can be composed "
                               "(the reference fails to broadcast depth / normals, rasterer.py:133-144)")
        if bg is not None and not output_nocs:
            output_nocs_eff = True        # rasterer.py:107-111 shows (colors + 1) / 2 whatever output_nocs says
        else:
            output_nocs_eff = bool(output_nocs)
        if rot not in ('dcm', 'quat'):
            raise ValueError(rot)
        if not coords.is_cuda:
            raise _lib.SdfrError("sdflabel_b200.Rasterer runs on a CUDA device only (no CPU path)")
        kinv, kmat = self._intrinsics()
        cfg = (int(self.res_x_px), int(self.res_y_px), kinv, kmat,
               _lib.ROT_DCM if rot == 'dcm' else _lib.ROT_QUAT, output_nocs_eff, prim, bg)
        color, mask, depth, nmap, xyz, rgb, xyzf, rgbf = _Splat.apply(coords, normals, colors, camera_matrix, cfg)
        rendering = {'color': color}
        if output_mask:
            rendering['mask'] = mask
        if output_depth:
            rendering['depth'] = depth
        if output_normals:
            rendering['normals'] = nmap
        if output_points:
            points = {'xyz': xyz, 'rgb': rgb}
            if rot == 'dcm':
                points['xyzf'] = xyzf
                points['rgbf'] = rgbf
            else:
                # the reference raises KeyError('points_3d_filt') here (rasterer.py:151 vs projection.py:160);
                # the unfiltered lists are returned instead so the quaternion path is usable
                points['xyzf'] = xyz
                points['rgbf'] = rgb
            return rendering, points
        return rendering
