"""Trace-mode renderer: per-ray sphere tracing of the DeepSDF decoder on the device.

This is the renderer BASELINE.json's north_star describes; the reference has no counterpart
(its renderer is the surfel splat, mirrored in ``rasterer.py``).  Same camera conventions as
``Rasterer``: ``K`` (3,3), resolution ``(W, H)``, ``camera_matrix`` a 4x4 object-to-camera DCM
pose with an orthogonal rotation (e.g. the refine loop's ``diag(1,-1,1) R_y(yaw)``, optimizer.py:87-90).

    tracer = SphereTracer(K, (W, H)).to(device)
    rendering = tracer(dsdf, latent, pose)      # dict: depth (1,H,W), normals (3,H,W), color = NOCS (3,H,W), mask

``depth`` and ``color`` are differentiable with respect to ``pose`` and ``latent`` (implicit
differentiation at the hit, ``sdfr_trace_backward``); ``normals`` and ``mask`` are constants,
like the normals in the reference's graph (grid.py:55-58).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib


class _Trace(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent_unit, pose, native, cfg_tuple, max_steps, eps, impl, cache, trusted=False, pose_h=None):
        lib = _lib.load()
        width, height, kinv, kmat = cfg_tuple
        dev = latent_unit.device
        cfg = _lib.RasterCfg(width=width, height=height, rot=_lib.ROT_DCM, output_nocs=1)
        cfg.kinv[:] = kinv
        cfg.k[:] = kmat
        lat = latent_unit.detach().contiguous().float()
        if pose_h is None:
            pose_h = np.ascontiguousarray(pose.detach().cpu().float().numpy().reshape(16))
        f32 = dict(device=dev, dtype=torch.float32)
        depth = torch.empty((1, height, width), **f32)
        nmap = torch.empty((3, height, width), **f32)
        nocs = torch.empty((3, height, width), **f32)
        mask = torch.empty((1, height, width), **f32)
        ws = torch.empty((lib.sdfr_trace_workspace_bytes(cfg, native.handle),), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(lib.sdfr_trace_forward(native.handle, cfg, lat.data_ptr(), _lib.fptr(pose_h), int(max_steps),
                                              float(eps), depth.data_ptr(), nmap.data_ptr(), nocs.data_ptr(),
                                              mask.data_ptr(), 0, ws.data_ptr(), _lib.ptr(cache),
                                              (-1.0 if trusted else float(native.latent_lipschitz)) if cache is not None
                                              else 0.0, impl, _lib.stream_ptr()))
        ctx.native, ctx.cfg, ctx.ws, ctx.pose_h, ctx.eps = native, cfg, ws, pose_h, float(eps)
        ctx.meta = (latent_unit.dtype, pose.dtype, pose.device, tuple(pose.shape), latent_unit.shape[0])
        ctx.mark_non_differentiable(nmap, mask)
        return depth, nmap, nocs, mask

    @staticmethod
    def backward(ctx, g_depth, _g_nmap, g_nocs, _g_mask):
        lib = _lib.load()
        lat_dtype, pose_dtype, pose_dev, pose_shape, L = ctx.meta
        dev = ctx.ws.device
        gd = g_depth.contiguous().float() if g_depth is not None else None
        gn = g_nocs.contiguous().float() if g_nocs is not None else None
        d_pose = torch.zeros(12, device=dev, dtype=torch.float32)
        d_lat = torch.zeros(L, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.sdfr_trace_backward(ctx.native.handle, ctx.cfg, _lib.fptr(ctx.pose_h), ctx.eps, _lib.ptr(gd),
                                               _lib.ptr(gn), d_pose.data_ptr(), d_lat.data_ptr(), ctx.ws.data_ptr(),
                                               _lib.stream_ptr()))
        g_pose = torch.zeros(pose_shape, device=dev, dtype=torch.float32)
        g_pose[:3, :4] = d_pose.view(3, 4)
        return d_lat.to(lat_dtype), g_pose.to(pose_dev, pose_dtype), None, None, None, None, None, None, None, None


class SphereTracer(torch.nn.Module):
    def __init__(self, K, resolution_px, max_steps=64, eps=1e-4, precision=torch.float32, reuse_cache=True):
        super().__init__()
        self.res_x_px, self.res_y_px = resolution_px
        self.max_steps, self.eps = int(max_steps), float(eps)
        self.register_buffer('K', K.to(precision))
        self._k_cache = None
        # distance cache of the fused march, kept across calls (sdfr_trace_cache_bytes): it depends on the latent
        # only, so rendering one shape from many poses - or a latent that moves by less than the decoder's
        # Lipschitz slack - pays for it once.  One block per native decoder handle.
        self.reuse_cache = bool(reuse_cache)
        self._dist_cache = None
        self._view_streams = []

    def _cache_for(self, native, device):
        if not self.reuse_cache:
            return None
        ent = self._dist_cache
        if ent is None or ent[0] is not native or ent[1].device != device:
            buf = torch.zeros((_lib.load().sdfr_trace_cache_bytes(),), device=device, dtype=torch.uint8)
            ent = self._dist_cache = (native, buf)
        return ent[1]

    def _intrinsics(self):
        k32 = self.K.detach().float().cpu()
        key = k32.numpy().tobytes()                       # by value: 36 bytes, never a stale pointer / version pair
        if self._k_cache is None or self._k_cache[0] != key:
            self._k_cache = (key, k32.inverse().reshape(-1).tolist(), k32.reshape(-1).tolist())
        return self._k_cache[1], self._k_cache[2]

    def forward(self, dsdf, latent, camera_matrix, normalize_latent=True):
        """latent: (L,) (normalised like the refine loop does, optimizer.py:96, unless told otherwise)."""
        if not latent.is_cuda:
            raise _lib.SdfrError("SphereTracer runs on a CUDA device only (no CPU path)")
        lat = torch.nn.functional.normalize(latent, p=2, dim=0) if normalize_latent else latent
        kinv, kmat = self._intrinsics()
        cfg = (int(self.res_x_px), int(self.res_y_px), kinv, kmat)
        native = dsdf.native()
        impl = getattr(dsdf, 'mlp_impl', _lib.MLP_AUTO)
        depth, nmap, nocs, mask = _Trace.apply(lat, camera_matrix, native, cfg, self.max_steps, self.eps, impl,
                                               self._cache_for(native, lat.device))
        return {'depth': depth, 'normals': nmap, 'color': nocs, 'mask': mask}

    def render_views(self, dsdf, latent, camera_matrices, normalize_latent=True, views_in_flight=4):
        """Forward-only renders of one latent from several poses, ``views_in_flight`` of them concurrently on
        separate CUDA streams.  A single trace is a chain of ~25 dependent launches, most of them bound by the latency
        of one decoder tile rather than by throughput (few rays are left after the first march steps and CTAs
        without rows exit at once), so independent views fill the SMs a single view leaves idle.  The views share ONE
        distance cache, brought up to date for the latent before they start (sdfr_trace_cache_update).  Pass the
        poses as HOST tensors: reading a device pose back synchronises its stream.
        Returns one dict per pose, ordered like ``camera_matrices``, usable on the current stream."""
        if not latent.is_cuda:
            raise _lib.SdfrError("SphereTracer runs on a CUDA device only (no CPU path)")
        dev = latent.device
        native = dsdf.native()
        kinv, kmat = self._intrinsics()
        cfg = (int(self.res_x_px), int(self.res_y_px), kinv, kmat)
        impl = getattr(dsdf, 'mlp_impl', _lib.MLP_AUTO)
        n_streams = max(1, min(int(views_in_flight), len(camera_matrices)))
        if not native.tcgen05 or impl == _lib.MLP_FFMA:
            n_streams = 1          # the CUDA-core decoder kernel spills to a scratch buffer the launches of a decoder share
        with torch.cuda.device(dev), torch.no_grad():
            while len(self._view_streams) < n_streams:
                self._view_streams.append(torch.cuda.Stream(device=dev))
            cur = torch.cuda.current_stream()
            lat = torch.nn.functional.normalize(latent, p=2, dim=0) if normalize_latent else latent
            lat = lat.detach().contiguous().float()
            fused = bool(native.tcgen05 and impl != _lib.MLP_FFMA)
            cache = None
            if fused:
                if self.reuse_cache:
                    cache, lip = self._cache_for(native, dev), float(native.latent_lipschitz)
                else:                        # rebuilt every call, still one block for all views of the call
                    cache = torch.zeros((_lib.load().sdfr_trace_cache_bytes(),), device=dev, dtype=torch.uint8)
                    lip = 0.0
                _lib.check(_lib.load().sdfr_trace_cache_update(native.handle, lat.data_ptr(), cache.data_ptr(), lip,
                                                               _lib.stream_ptr()))
            ready = torch.cuda.Event()
            ready.record(cur)
            outs = []
            for i, pose in enumerate(camera_matrices):
                st = self._view_streams[i % n_streams]
                if i < n_streams:
                    st.wait_event(ready)
                with torch.cuda.stream(st):
                    maps = _Trace.apply(lat, pose, native, cfg, self.max_steps, self.eps, impl, cache, fused)
                for t in maps:
                    t.record_stream(cur)
                outs.append(dict(zip(('depth', 'normals', 'color', 'mask'), maps)))
            for st in self._view_streams[:n_streams]:
                lat.record_stream(st)
                if cache is not None:
                    cache.record_stream(st)
                cur.wait_stream(st)
        return outs
