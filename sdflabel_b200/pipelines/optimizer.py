"""Pose / shape refinement loop on the device.

Mirror of the reference's pipelines/optimizer.py: ``MultipleOptimizer`` (13-23),
``get_opt_params`` (26-40), ``Optimizer.__init__`` (43-54) and
``Optimizer.optimize`` (56-164) with the same signatures; results are read back
from the same ``params`` dict, mutated in place to tensors, exactly as
``refine_css.py:229-231`` expects.

The loop body does not run in PyTorch.  ``optimize`` hands the host inputs to the
fused engine of libsdfr.so (``sdfr_refine_*``): per iteration the DeepSDF lattice
evaluation with its input gradient, the band extraction, the surfel splat, both
losses, every gradient and the Adam/SGD update are CUDA kernels chained on one
stream without host synchronisation (the reference does one H2D, one D2H + a CPU
KD-tree and two ``.item()`` syncs per iteration).  ``BatchOptimizer`` refines all
detections of a frame (or many frames) in the same launches.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib


class MultipleOptimizer:
    def __init__(self, *op):
        self.optimizers = op

    def zero_grad(self):
        for op in self.optimizers:
            op.zero_grad()

    def step(self):
        for op in self.optimizers:
            op.step()


def get_opt_params(params, device):
    """numpy -> leaf tensors and the four parameter groups with the reference's
    learning rates (optimizer.py:26-40).  The fused engine applies the same rates."""
    for key, value in params.items():
        if isinstance(value, torch.Tensor):
            params[key] = value.detach().to(device, torch.float32).requires_grad_(True)
        else:
            params[key] = torch.tensor(np.asarray(value, dtype=np.float32)).to(device).requires_grad_(True)
    optim_params = [
        {'params': params['yaw'], 'lr': 0.01},
        {'params': params['trans'], 'lr': 0.01},
        {'params': params['scale'], 'lr': 0.01},
        {'params': params['latent'], 'lr': 0.00003},
    ]
    return params, optim_params


def _pow2_ceil(v: int, lo: int) -> int:
    v = max(int(v), 1)
    return max(lo, 1 << (v - 1).bit_length())


class _Engine:
    """One sdfr_refine handle: ``batch`` detection slots with fixed crop / LIDAR / history capacities,
    of which a run covers the first ``active``.  Cached on the decoder object and only ever re-created
    to GROW a capacity, so a sequence of frames with different detection counts and crop sizes
    allocates (and captures its CUDA graphs) once."""

    def __init__(self, native_decoder, batch, density, max_w, max_h, max_lidar, max_iters, w2d, w3d, impl,
                 latent_lipschitz=0.0):
        lib = _lib.load()
        self.cfg = _lib.RefineCfg(batch=batch, density=density, max_width=max_w, max_height=max_h,
                                  max_lidar=max_lidar, max_iters=max_iters, weight_2d=w2d, weight_3d=w3d,
                                  mlp_impl=impl, latent_lipschitz=float(latent_lipschitz))
        self.native_decoder = native_decoder     # keeps the decoder handle alive
        self.latent_size = native_decoder.latent_size
        h = _lib.vp()
        _lib.check(lib.sdfr_refine_create(native_decoder.handle, C.byref(self.cfg), C.byref(h)))
        self.handle = h
        self.active = batch
        self._pinned: Dict = {}     # persistent pinned staging buffers, keyed by (name, detection)
        self._dirty = set()         # detections whose staging buffers may still be in flight
        self._kcache: Dict = {}
        self._out = None            # read-back arrays of optimize_slot

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().sdfr_refine_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def capacity(self):
        c = self.cfg
        return (c.batch, c.max_width, c.max_height, c.max_lidar, c.max_iters)

    def _pin(self, key, arr) -> torch.Tensor:
        """float32 host copy in a persistent pinned buffer, so the H2D copy is a true async DMA and
        no cudaHostAlloc happens per call.  The copy is a plain single-threaded numpy memcpy: torch's
        CPU copy_ goes through the OpenMP pool above 32 K elements, and waking a sleeping pool on a
        128-core host costs tens of milliseconds every few calls (measured)."""
        if isinstance(arr, torch.Tensor):
            src = arr.detach()
            if src.device.type == 'cpu' and src.dtype == torch.float32 and src.is_contiguous() and src.is_pinned():
                return src                       # already page-locked: DMA straight from the caller's buffer
            src = (src if src.device.type == 'cpu' else src.cpu()).numpy()
        else:
            src = np.asarray(arr)
        n = int(src.size)
        ent = self._pinned.get(key)
        if ent is None or ent[0].numel() < n:
            buf = torch.empty((max(n, 1),), dtype=torch.float32, pin_memory=True)
            ent = (buf, buf.numpy())
            self._pinned[key] = ent
        np.copyto(ent[1][:n], src.reshape(-1), casting='same_kind')
        return ent[0][:n].view(tuple(src.shape))

    def _intrinsics(self, K):
        """(K, K^-1) as contiguous fp32 host tensors; the inverse is torch's own K.float().inverse()
        (primitives.py:204).  Cached by VALUE (the 36 bytes of K): a fresh K tensor per detection usually
        lands in the allocation the previous one just freed, so nothing about the tensor object identifies it."""
        if isinstance(K, torch.Tensor):
            k32 = K.detach()
            if k32.device.type != 'cpu' or k32.dtype != torch.float32 or not k32.is_contiguous():
                k32 = k32.float().cpu().contiguous()
        else:
            k32 = torch.from_numpy(np.ascontiguousarray(np.asarray(K, dtype=np.float32)))
        key = k32.numpy().tobytes()
        hit = self._kcache.get(key)
        if hit is None:
            if len(self._kcache) > 256:
                self._kcache.clear()
            hit = (k32.clone(), k32.inverse().contiguous())
            self._kcache[key] = hit
        return hit

    def set_detection(self, b, K, width, height, nocs_pred, lidar_np, yaw=None, trans=None, scale=None, latent=None):
        """Inputs of slot b.  The parameters are host arrays, or all None when ``import_params`` follows."""
        lib = _lib.load()
        k32, kinv = self._intrinsics(K)
        # the previous use of these staging buffers must have been consumed by the device
        if b in self._dirty:
            torch.cuda.current_stream().synchronize()
            self._dirty.clear()
        nocs = self._pin(('nocs', b), nocs_pred)
        lidar = self._pin(('lidar', b), np.asarray(lidar_np, dtype=np.float32).reshape(-1, 3))
        fp = lambda t: C.cast(t.data_ptr(), _lib.c_float_p)
        par = [None if v is None else fp(self._pin((n, b), np.asarray(v, dtype=np.float32).reshape(-1)))
               for n, v in (('yaw', yaw), ('trans', trans), ('scale', scale), ('latent', latent))]
        self._dirty.add(b)
        _lib.check(lib.sdfr_refine_set_detection(
            self.handle, b, fp(k32), fp(kinv), int(width), int(height), nocs.data_ptr(), int(nocs.shape[1]),
            int(nocs.shape[2]), lidar.data_ptr(), int(lidar.shape[0]), par[0], par[1], par[2], par[3],
            _lib.stream_ptr()))

    def import_params(self, b, p):
        """Initial parameters of slot b straight from the caller's device tensors (no host round trip)."""
        _lib.check(_lib.load().sdfr_refine_import(self.handle, b, p['yaw'].data_ptr(), p['trans'].data_ptr(),
                                                  p['scale'].data_ptr(), p['latent'].data_ptr(), _lib.stream_ptr()))

    def export_params(self, b, p):
        _lib.check(_lib.load().sdfr_refine_export(self.handle, b, p['yaw'].data_ptr(), p['trans'].data_ptr(),
                                                  p['scale'].data_ptr(), p['latent'].data_ptr(), _lib.stream_ptr()))

    def set_active(self, n):
        _lib.check(_lib.load().sdfr_refine_set_active(self.handle, int(n)))
        self.active = int(n)

    @staticmethod
    def _host_f32(arr):
        """contiguous float32 host array sharing memory with ``arr`` when it already is one (tensor or ndarray)."""
        if isinstance(arr, torch.Tensor):
            arr = arr.detach()
            arr = (arr if arr.device.type == 'cpu' else arr.cpu()).numpy()
        return np.ascontiguousarray(arr, dtype=np.float32)

    def optimize_slot(self, b, K, width, height, nocs_pred, lidar_np, p, adam, iters):
        """``Optimizer.optimize`` of slot b as ONE C-ABI call (sdfr_refine_optimize): host inputs in, the device
        tensors of ``p`` read and updated in place, ``adam`` = (m[4], v[4], c_int t) continued and updated, one
        stream synchronisation.  The call is synchronous, so the host arrays are handed over as they are (page-locked
        or not) without staging copies."""
        k32, kinv = self._intrinsics(K)
        nocs = self._host_f32(nocs_pred)
        lidar = self._host_f32(lidar_np).reshape(-1, 3)
        out = self._out
        if out is None or out[1].shape[0] < self.cfg.max_iters:
            out = self._out = (np.zeros(5 + self.latent_size, dtype=np.float32),
                               np.zeros((self.cfg.max_iters, 4), dtype=np.float32), C.c_int(0))
        if self._dirty:                      # staged copies of an earlier set_detection may still be in flight
            torch.cuda.current_stream().synchronize()
            self._dirty.clear()
        _lib.check(_lib.load().sdfr_refine_optimize(
            self.handle, b, k32.data_ptr(), kinv.data_ptr(), int(width), int(height), nocs.ctypes.data,
            int(nocs.shape[1]), int(nocs.shape[2]), lidar.ctypes.data, int(lidar.shape[0]), p['yaw'].data_ptr(),
            p['trans'].data_ptr(), p['scale'].data_ptr(), p['latent'].data_ptr(), adam[0].ctypes.data,
            adam[1].ctypes.data, C.addressof(adam[2]), int(iters), out[0].ctypes.data, out[1].ctypes.data,
            C.addressof(out[2]), _lib.stream_ptr()))
        return out[0].copy(), out[1][:out[2].value].copy()

    def set_optimizer_state(self, b, state):
        m, v, t = state
        m = np.ascontiguousarray(m, dtype=np.float32)
        v = np.ascontiguousarray(v, dtype=np.float32)
        _lib.check(_lib.load().sdfr_refine_set_optimizer_state(self.handle, b, _lib.fptr(m), _lib.fptr(v), int(t),
                                                               _lib.stream_ptr()))

    def get_optimizer_state(self, b):
        """(adam_m[4], adam_v[4], adam_t) of slot b as of the last ``get`` / ``get_batch``."""
        m = np.zeros(4, dtype=np.float32)
        v = np.zeros(4, dtype=np.float32)
        t = C.c_int(0)
        _lib.check(_lib.load().sdfr_refine_get_optimizer_state(self.handle, b, _lib.fptr(m), _lib.fptr(v), C.byref(t)))
        return m, v, t.value

    def run(self, iters):
        _lib.check(_lib.load().sdfr_refine_run(self.handle, int(iters), _lib.stream_ptr()))

    def get(self, b):
        L = self.latent_size
        params = np.zeros(5 + L, dtype=np.float32)
        hist = np.zeros((self.cfg.max_iters, 4), dtype=np.float32)
        nh = C.c_int(0)
        try:
            _lib.check(_lib.load().sdfr_refine_get(self.handle, b, _lib.fptr(params), _lib.fptr(hist), C.byref(nh),
                                                   _lib.stream_ptr()))
        finally:
            self._dirty.clear()     # sdfr_refine_get synchronised the stream: all staged copies are done
        return params, hist[:nh.value]

    def get_batch(self):
        """(params [active, 5+L], [history rows of each active detection]) in one synchronisation."""
        L, B = self.latent_size, self.active
        params = np.zeros((B, 5 + L), dtype=np.float32)
        hist = np.zeros((B, self.cfg.max_iters, 4), dtype=np.float32)
        nh = (C.c_int * B)()
        try:
            _lib.check(_lib.load().sdfr_refine_get_batch(self.handle, _lib.fptr(params), _lib.fptr(hist), nh,
                                                         _lib.stream_ptr()))
        finally:
            self._dirty.clear()
        return params, [hist[b, :nh[b]] for b in range(B)]

    def lattice_rows(self, reset=False):
        """(lattice points evaluated by the pruned lattice passes, detection-iterations served) since the last reset;
        (0, 0) for an engine without temporal pruning."""
        rows, its = C.c_int64(0), C.c_int64(0)
        _lib.check(_lib.load().sdfr_refine_lattice_rows(self.handle, C.byref(rows), C.byref(its), int(bool(reset)),
                                                        _lib.stream_ptr()))
        return rows.value, its.value

    def preselect_error(self) -> float:
        e = np.zeros(1, dtype=np.float32)
        _lib.check(_lib.load().sdfr_refine_preselect_error(self.handle, _lib.fptr(e), _lib.stream_ptr()))
        return float(e[0])

    def surface_clouds(self, latents):
        """Isosurface points (device tensors (M_b, 3), object frame) of the raw latents [n, L]: the model clouds of
        the pose initialisation (refine_css.py:143-151), all n in one lattice pass on slots [0, n)."""
        lat = np.ascontiguousarray(latents, dtype=np.float32).reshape(-1, self.latent_size)
        n = lat.shape[0]
        lib = _lib.load()
        pin = self._pin(('latents', 0), lat)
        for b in range(n):
            _lib.check(lib.sdfr_refine_set_latent(self.handle, b, C.cast(pin.data_ptr() + 4 * b * self.latent_size,
                                                                       _lib.c_float_p), _lib.stream_ptr()))
        if self.active != n:
            self.set_active(n)
        ext = self.label_extents()
        clouds = []
        for b in range(n):
            m = int(ext[b, 7])
            keep = self.view(b, 'surf_valid', m).bool()
            clouds.append(self.view(b, 'surf_pts', 3 * m).view(-1, 3)[keep])
        return clouds

    def label_extents(self):
        """[active, 8]: min xyz, max xyz of the isosurface points of the CURRENT raw latent, band point count,
        pre-selected row count (the extents ``get_kitti_label`` derives, utils/refinement.py:527-541)."""
        out = np.zeros((self.active, 8), dtype=np.float32)
        _lib.check(_lib.load().sdfr_refine_label_extents(self.handle, _lib.fptr(out), _lib.stream_ptr()))
        return out

    VIEW_KINDS = {'sdf': 0, 'dinput': 1, 'surf_pts': 2, 'surf_nrm': 3, 'color': 4, 'mask': 5, 'normals': 6,
                  'grads': 7, 'surf_count': 8, 'depth': 9, 'cam_pts': 10, 'front': 11, 'surf_valid': 12, 'cam_rgb': 13,
                  'surf_idx': 14}

    def view(self, b, kind, count=None):
        """Device copy of an intermediate of the last iteration (tests, label dumps); the first ``count``
        elements when given."""
        lib = _lib.load()
        kind = self.VIEW_KINDS[kind] if isinstance(kind, str) else int(kind)
        p = _lib.vp()
        n = C.c_int64(0)
        _lib.check(lib.sdfr_refine_view(self.handle, b, kind, C.byref(p), C.byref(n)))
        size = n.value if count is None else min(int(count), n.value)
        dtype = torch.int32 if kind in (8, 14) else torch.uint8 if kind in (11, 12) else torch.float32
        out = torch.empty((size,), device='cuda', dtype=dtype)
        _lib.check(lib.sdfr_refine_copy_view(self.handle, b, kind, out.data_ptr(), size, _lib.stream_ptr()))
        return out

    def surfels(self, b):
        """(points (M,3), normals (M,3)) of the band of detection b after the last iteration."""
        m = int(self.view(b, 'surf_count').item())
        keep = self.view(b, 'surf_valid')[:m].bool()
        return self.view(b, 'surf_pts')[:m * 3].view(-1, 3)[keep], self.view(b, 'surf_nrm')[:m * 3].view(-1, 3)[keep]

    def front_points(self, b):
        """points['xyzf'] / points['rgbf'] of the last iteration (projection.py:61-70, rasterer.py:151-152)."""
        m = int(self.view(b, 'surf_count').item())
        keep = self.view(b, 'front')[:m].bool()
        return self.view(b, 'cam_pts')[:m * 3].view(-1, 3)[keep], self.view(b, 'cam_rgb')[:m * 3].view(-1, 3)[keep]


# Temporal pruning of the lattice pass (sdfr_refine_cfg.latent_lipschitz): an iteration evaluates only the lattice
# points the decoder's Lipschitz bound cannot exclude from the band.  Same surfels and results as evaluating the
# whole lattice every iteration (tests/test_gpu_parity.py::test_temporal_pruning_is_exact); False restores that.
TEMPORAL_PRUNING = True


def _engine_for(dsdf, batch, density, w, h, n_lidar, iters, weights, impl, prune=None) -> _Engine:
    """The decoder's engine for this (density, loss weights, MLP kernel), grown when a request exceeds a
    capacity.  Capacities are rounded up to powers of two, so growth happens a handful of times at most;
    nothing is re-created when a frame simply has fewer detections or a smaller crop than the last one."""
    native = dsdf.native()
    prune = TEMPORAL_PRUNING if prune is None else bool(prune)
    lip = float(native.latent_lipschitz) if prune else 0.0
    cache = dsdf.__dict__.setdefault('_sdfr_engines', {})
    for k in [k for k, e in cache.items() if e.native_decoder is not native]:
        del cache[k]           # engines of a decoder handle that has been rebuilt (weights changed)
    need = (_pow2_ceil(batch, 1), _pow2_ceil(w, 32), _pow2_ceil(h, 32), _pow2_ceil(max(n_lidar, 1), 1024),
            max(64, int(iters)))
    key = (int(density), float(weights['2d']), float(weights['3d']), int(impl), lip)
    eng = cache.get(key)
    if eng is None or any(n > c for n, c in zip(need, eng.capacity())):
        cap = need if eng is None else tuple(max(n, c) for n, c in zip(need, eng.capacity()))
        if eng is not None:
            del cache[key]
            eng = None         # frees the smaller engine's device memory before the larger one allocates
        eng = _Engine(native, cap[0], int(density), cap[1], cap[2], cap[3], cap[4], key[1], key[2], key[3], lip)
        cache[key] = eng
    return eng


class _Loss3D(torch.autograd.Function):
    """mean ||L_nn - v|| over pairs closer than `radius` (sdfr_loss3d: exact brute-force 1-NN on the device)."""

    @staticmethod
    def forward(ctx, xyzf, lidar_scaled, radius):
        lib = _lib.load()
        q = xyzf.detach().contiguous().float()
        l = lidar_scaled.detach().contiguous().float()
        loss = torch.zeros(2, device=q.device)
        dq = torch.zeros_like(q)
        dl = torch.zeros_like(l)
        with torch.cuda.device(q.device):
            _lib.check(lib.sdfr_loss3d(q.data_ptr(), q.shape[0], l.data_ptr(), l.shape[0], float(radius),
                                       loss.data_ptr(), dq.data_ptr(), dl.data_ptr(), _lib.stream_ptr()))
        ctx.save_for_backward(dq, dl)
        ctx.dtypes = (xyzf.dtype, lidar_scaled.dtype)
        return loss[0].to(xyzf.dtype)

    @staticmethod
    def backward(ctx, g):
        dq, dl = ctx.saved_tensors
        return (g * dq).to(ctx.dtypes[0]), (g * dl).to(ctx.dtypes[1]), None


class _Loss2D(torch.autograd.Function):
    """Windowed form of the reference's O(M*H*W) NOCS loss (sdfr_loss2d)."""

    @staticmethod
    def forward(ctx, color, target):
        lib = _lib.load()
        c = color.detach().contiguous().float()
        t = target.detach().contiguous().float()
        loss = torch.zeros(2, device=c.device)
        dc = torch.zeros_like(c)
        with torch.cuda.device(c.device):
            _lib.check(lib.sdfr_loss2d(c.data_ptr(), t.data_ptr(), c.shape[1], c.shape[2], loss.data_ptr(),
                                       dc.data_ptr(), _lib.stream_ptr()))
        ctx.save_for_backward(dc)
        ctx.dtype = color.dtype
        return loss[0].to(color.dtype)

    @staticmethod
    def backward(ctx, g):
        (dc,) = ctx.saved_tensors
        return (g * dc).to(ctx.dtype), None


class Optimizer:
    def __init__(self, params, device, weights, rot='dcm'):
        self.params, self.optim_params = get_opt_params(params, device)
        self.optim_params_adam = self.optim_params[:2]
        self.optim_params_sgd = self.optim_params[2:]
        self.weights = weights
        self.rot = rot
        self.verbose = False
        self.history = None
        # Adam moments / step count of (yaw, trans): the reference builds its solver once per Optimizer
        # (optimizer.py:46-52), so n calls of optimize(k) continue one state exactly like optimize(n*k)
        self._adam = None

    def optimize(self, iters_optim, nocs_pred, pcd_frustum_np, dsdf, grid, K, crop_size, viz_type=None,
                 frame_vis=None):
        """
        Optimization loop (same arguments as the reference).
        Args:
            iters_optim (int): number of iterations
            nocs_pred (torch.Tensor): CSS network prediction (3,h,w)
            pcd_frustum_np (np.array): LIDAR point cloud (N,3)
            dsdf: sdflabel_b200 Decoder
            grid: sdflabel_b200 Grid3D
            K (torch.Tensor): camera matrix (3,3) of the crop
            crop_size (list): [H, W] of the optimized crop
            viz_type: must be None (open3d / cv2 visualisation is out of scope)
        """
        if viz_type not in (None, 'none'):
            raise NotImplementedError("visualisation is not part of the device refine loop")
        if self.rot != 'dcm':
            raise NotImplementedError("the refine loop optimises yaw/translation (rot='dcm'), as the reference does")
        self.device = grid.points.device
        self.precision = grid.points.dtype
        height, width = int(crop_size[0]), int(crop_size[1])
        lidar = np.asarray(pcd_frustum_np, dtype=np.float32).reshape(-1, 3)
        p = self.params
        keys = ('yaw', 'trans', 'scale', 'latent')
        with torch.cuda.device(self.device):
            eng = _engine_for(dsdf, 1, int(grid.density), width, height, lidar.shape[0], iters_optim, self.weights,
                              getattr(dsdf, 'mlp_impl', _lib.MLP_AUTO))
            # the params tensors are read and updated in place ON THE DEVICE (one launch each way): whatever
            # the caller did to them since the last call (in-place edits included) is what the loop starts from
            direct = all(p[k].is_cuda and p[k].dtype == torch.float32 and p[k].is_contiguous() and
                         p[k].device == self.device for k in keys)
            if eng.active != 1:
                eng.set_active(1)
            if direct:
                # one C-ABI call: inputs, parameters in / out on the device, Adam state, iterations, read-back
                if self._adam is None:
                    self._adam = (np.zeros(4, dtype=np.float32), np.zeros(4, dtype=np.float32), C.c_int(0))
                out, hist = eng.optimize_slot(0, K, width, height, nocs_pred, lidar, p, self._adam, iters_optim)
            else:
                host = {k: p[k].detach().cpu().numpy() for k in keys}
                eng.set_detection(0, K, width, height, nocs_pred, lidar, host['yaw'], host['trans'], host['scale'],
                                  host['latent'])
                if self._adam is not None:
                    eng.set_optimizer_state(0, (self._adam[0], self._adam[1], self._adam[2].value))
                eng.run(iters_optim)
                # read back (one stream sync; it also carries the decoder's fp16-range and pre-selection guards)
                out, hist = eng.get(0)
                m, v, t = eng.get_optimizer_state(0)
                self._adam = (m, v, C.c_int(t))
        self.engine = eng
        self.history = hist
        if not direct:
            new = {'yaw': out[0:1], 'trans': out[1:4], 'scale': out[4:5], 'latent': out[5:]}
            with torch.no_grad():
                for k, v in new.items():
                    p[k].copy_(torch.from_numpy(v.copy()).to(p[k].dtype))
        if self.verbose:
            w2, w3 = self.weights['2d'], self.weights['3d']
            for e, (l2, l3, tot, skip) in enumerate(hist):
                if skip:
                    print('Skip frame')
                else:
                    print('ITER {} | Losses: 2D - {}, 3D - {}, Total - {}'.format(e, w2 * l2, w3 * l3, tot))

    # ---- stand-alone losses with the reference's signatures (optimizer.py:166-237) -------------------
    def compute_loss_3d(self, pcd_dsdf_trans, pcd_frustum, threshold=0.2):
        """3D loss between the estimated and LIDAR point clouds; returns (loss, None, None): the
        neighbour distances / indices the reference also returns only feed its open3d visualiser."""
        if pcd_dsdf_trans.nelement() == 0 or pcd_frustum.nelement() == 0:
            return torch.zeros((), device=pcd_dsdf_trans.device, dtype=pcd_dsdf_trans.dtype), None, None
        radius = threshold / self.params['scale'][0].item()
        return _Loss3D.apply(pcd_dsdf_trans, pcd_frustum, radius), None, None

    def compute_loss_2d(self, rendering_nocs, css_nocs, diam=5, threshold_nocs=1):
        """2D loss between the CSS net output and the rendering (diam / threshold are the reference's
        fixed 5 px / 1.0; other values are not supported by the kernel)."""
        if diam != 5 or threshold_nocs != 1:
            raise NotImplementedError("sdfr_loss2d implements the reference defaults diam=5, threshold_nocs=1")
        return _Loss2D.apply(rendering_nocs, css_nocs.to(rendering_nocs.device))


class BatchOptimizer:
    """All detections of one or more frames refined in the same kernel launches.

    ``detections`` is a sequence of dicts with the arguments ``Optimizer`` takes
    per detection: ``params`` (yaw/trans/scale/latent), ``nocs_pred``, ``lidar``,
    ``K``, ``crop_size``.  This is the per-GPU unit of the multi-GPU pipeline
    (frames are sharded across ranks).  Every detection starts
    from a fresh optimiser state, like the ``Optimizer`` the reference constructs per
    annotation (refine_css.py:203)."""

    def __init__(self, weights, device='cuda'):
        self.weights = weights
        self.device = torch.device(device)

    def reserve(self, dsdf, grid, max_batch, max_crop, max_lidar=1024, iters=64):
        """Allocates the engine for the largest batch / crop / LIDAR crop up front."""
        with torch.cuda.device(self.device):
            self.engine = _engine_for(dsdf, max_batch, int(grid.density), max_crop[1], max_crop[0], max_lidar, iters,
                                      self.weights, getattr(dsdf, 'mlp_impl', _lib.MLP_AUTO))
        return self.engine

    def optimize(self, iters_optim, detections: Sequence[Dict], dsdf, grid, extents=False):
        B = len(detections)
        if B == 0:
            return []
        max_w = max(int(d['crop_size'][1]) for d in detections)
        max_h = max(int(d['crop_size'][0]) for d in detections)
        max_l = max(int(np.asarray(d['lidar']).reshape(-1, 3).shape[0]) for d in detections)
        with torch.cuda.device(self.device):
            eng = _engine_for(dsdf, B, int(grid.density), max_w, max_h, max_l, iters_optim, self.weights,
                              getattr(dsdf, 'mlp_impl', _lib.MLP_AUTO))
            if eng.active != B:
                eng.set_active(B)
            for b, d in enumerate(detections):
                p = d['params']
                eng.set_detection(b, d['K'], int(d['crop_size'][1]), int(d['crop_size'][0]), d['nocs_pred'],
                                  d['lidar'], p['yaw'], p['trans'], p['scale'], p['latent'])
            eng.run(iters_optim)
            out, hists = eng.get_batch()
            ext = eng.label_extents() if extents else None
        results = []
        for b in range(B):
            res = {'yaw': out[b, 0:1].copy(), 'trans': out[b, 1:4].copy(), 'scale': out[b, 4:5].copy(),
                   'latent': out[b, 5:].copy(), 'history': hists[b].copy()}
            if ext is not None:
                res['extent_min'], res['extent_max'], res['extent_count'] = ext[b, 0:3].copy(), ext[b, 3:6].copy(), int(ext[b, 6])
            results.append(res)
        self.engine = eng
        return results
