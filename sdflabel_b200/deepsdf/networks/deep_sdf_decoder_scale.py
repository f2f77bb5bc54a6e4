"""DeepSDF decoder evaluated by libsdfr.so.

Mirror of ``Decoder`` in the reference's
sdfrenderer/deepsdf/networks/deep_sdf_decoder_scale.py:10-114: same constructor
arguments, same parameter names (``lin{l}.weight_g/weight_v/bias`` or ``.weight``,
``bn{l}``, ``scale_net``) so reference checkpoints load unchanged, same
``forward(input (N, L+3)) -> (sdf (N,1), scale)`` contract.  The MLP itself never
runs in PyTorch: ``forward`` hands the folded weights to ``sdfr_decoder_create``
once and calls ``sdfr_decoder_eval`` (tcgen05 kernel for the stock spec class,
fp32 CUDA-core kernel otherwise).  The input gradient that the reference gets
from ``pred_sdf_grid.sum().backward()`` (grid.py:55) comes out of the same launch.

Differences, by design: eval mode only (dropout is a no-op there, decoder.py:103);
gradients are produced for the *input* only - the reference also fills ``.grad``
of every decoder weight, which nothing on the refine path reads.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from ... import _lib


def spectral_norm_upper(w: np.ndarray, squarings: int = 9) -> float:
    """Upper bound of ||w||_2 that converges from ABOVE: for the Gram matrix A = w w^T (symmetric PSD),
    ||A||_2 = ||A^k||_2^(1/k) <= ||A^k||_F^(1/k); A^k by repeated squaring (k = 2^squarings: within
    rank^(1/2k) of the true value, 0.6 % for rank 512 and k = 512), float64, 1e-6 added for the rounding."""
    w = np.asarray(w, dtype=np.float64)
    if w.size == 0:
        return 0.0
    a = w @ w.T if w.shape[0] <= w.shape[1] else w.T @ w
    log_s = 0.0                                   # A^(2^i) = exp(log_s) * a
    for _ in range(squarings):
        f = np.linalg.norm(a)
        if not np.isfinite(f) or f == 0.0:
            return 0.0 if f == 0.0 else float('inf')
        a = (a / f) @ (a / f)
        log_s = 2.0 * (log_s + np.log(f))
    f = np.linalg.norm(a)
    if f == 0.0:
        return 0.0
    return float(np.exp((log_s + np.log(f)) / (2.0 ** squarings) / 2.0) * (1.0 + 1e-6))


def latent_lipschitz_bound(weights, concat, layer_norm, latent_size, in0) -> float:
    """Certified upper bound of |d sdf / d latent| (Euclidean) of Decoder.forward
    (deep_sdf_decoder_scale.py:78-114): ReLU, tanh and eval-mode dropout are 1-Lipschitz, a Linear contributes
    its spectral norm, and a layer that concatenates the input again adds the norm of its latent columns.
    0.0 (= no bound) for LayerNorm decoders."""
    if any(layer_norm):
        return 0.0
    g = 0.0
    for l, w in enumerate(weights):
        w = np.asarray(w.detach().cpu().numpy() if hasattr(w, 'detach') else w, dtype=np.float64)
        if l == 0:
            g = spectral_norm_upper(w[:, :latent_size])
        elif concat[l] == 1:                      # cat([x, latent, xyz])
            nh = w.shape[1] - in0
            g = spectral_norm_upper(w[:, :nh]) * g + spectral_norm_upper(w[:, nh:nh + latent_size])
        elif concat[l] == 2:                      # cat([x, xyz])
            g = spectral_norm_upper(w[:, :w.shape[1] - 3]) * g
        else:
            g = spectral_norm_upper(w) * g
    return float(g) if np.isfinite(g) else 0.0


class _NativeDecoder:
    """Owns one sdfr_decoder handle (device-resident folded weights)."""

    def __init__(self, layer_dims, concat, layer_norm, latent_size, use_tanh, weights, biases, ln_w, ln_b):
        lib = _lib.load()
        _lib.require_cuda()
        n = len(layer_dims)
        self._keep = []

        def iarr(vals):
            a = (C.c_int32 * n)(*[int(v) for v in vals])
            self._keep.append(a)
            return a

        def farrs(tensors):
            ptrs = (_lib.c_float_p * n)()
            for i, t in enumerate(tensors):
                if t is None:
                    ptrs[i] = None
                else:
                    a = np.ascontiguousarray(t.detach().cpu().float().numpy())
                    self._keep.append(a)
                    ptrs[i] = a.ctypes.data_as(_lib.c_float_p)
            self._keep.append(ptrs)
            return ptrs

        spec = _lib.DecoderSpec(
            latent_size=latent_size, num_layers=n,
            in_dims=iarr([d[0] for d in layer_dims]), out_dims=iarr([d[1] for d in layer_dims]),
            concat=iarr(concat), layer_norm=iarr(layer_norm), use_tanh=int(bool(use_tanh)))
        handle = _lib.vp()
        _lib.check(lib.sdfr_decoder_create(C.byref(spec), farrs(weights), farrs(biases), farrs(ln_w), farrs(ln_b),
                                           C.byref(handle)))
        self.handle = handle
        self.latent_size = latent_size
        self.tcgen05 = bool(lib.sdfr_decoder_tcgen05_ok(handle))
        self._keep = []   # the library copied everything
        self._lipschitz_args = (weights, list(concat), list(layer_norm), latent_size, layer_dims[0][0])
        self._lipschitz = None

    @property
    def latent_lipschitz(self) -> float:
        """Certified bound of |d sdf / d latent| (0.0 = none), computed on first use (0.1 - 0.5 s of numpy)."""
        if self._lipschitz is None:
            self._lipschitz = latent_lipschitz_bound(*self._lipschitz_args)
        return self._lipschitz

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().sdfr_decoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class _DecoderEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, native, impl):
        lib = _lib.load()
        x = inputs.detach().contiguous().float()
        n = x.shape[0]
        sdf = torch.empty((n, 1), device=x.device, dtype=torch.float32)
        need_grad = inputs.requires_grad
        dinput = torch.empty_like(x) if need_grad else None
        with torch.cuda.device(x.device):
            _lib.check(lib.sdfr_decoder_eval(native.handle, x.data_ptr(), n, sdf.data_ptr(), _lib.ptr(dinput), impl,
                                             _lib.stream_ptr()))
        ctx.save_for_backward(dinput) if need_grad else None
        ctx.in_dtype = inputs.dtype
        return sdf.to(inputs.dtype)

    @staticmethod
    def backward(ctx, grad_sdf):
        (dinput,) = ctx.saved_tensors
        return (grad_sdf.float() * dinput).to(ctx.in_dtype), None, None


class Decoder(nn.Module):
    def __init__(
        self,
        latent_size,
        dims,
        dropout=None,
        dropout_prob=0.0,
        norm_layers=(),
        latent_in=(),
        weight_norm=False,
        xyz_in_all=None,
        use_tanh=False,
        latent_dropout=False,
        samples_per_scene=None,
    ):
        super().__init__()
        full = [latent_size + 3] + list(dims) + [1]
        self.num_layers = len(full)
        self.latent_size = latent_size
        self.norm_layers = norm_layers
        self.latent_in = latent_in
        self.latent_dropout = latent_dropout
        self.xyz_in_all = xyz_in_all
        self.weight_norm = weight_norm
        self.samples_per_scene = samples_per_scene
        self.use_tanh = use_tanh
        self.dropout = dropout
        self.dropout_prob = dropout_prob
        self.mlp_impl = _lib.MLP_AUTO

        self._layer_dims = []
        for l in range(self.num_layers - 1):
            if l + 1 in latent_in:
                out_dim = full[l + 1] - full[0]
            else:
                out_dim = full[l + 1]
                if xyz_in_all and l != self.num_layers - 2:
                    out_dim -= 3
            lin = nn.Linear(full[l], out_dim)
            if weight_norm and l in norm_layers:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
            if (not weight_norm) and norm_layers is not None and l in norm_layers:
                setattr(self, "bn" + str(l), nn.LayerNorm(out_dim))
            self._layer_dims.append((full[l], out_dim))

        # same tiny scale head as the reference (decoder.py:69-75); evaluated with torch, off the hot path
        self.scale_net = nn.Sequential(
            nn.Linear(latent_size, 3), nn.ReLU(True), nn.Linear(3, 3), nn.ReLU(True), nn.Linear(3, 1))
        self._native = None
        self._native_key = None

    # ---- native handle ---------------------------------------------------------------------
    # The folded weights live in libsdfr's device memory; the handle is rebuilt whenever the parameters may
    # have changed: load_state_dict / .to() / .float() / ... invalidate it explicitly, in-place edits through
    # autograd-visible ops are caught by the tensors' version counters.  Raw ``.data`` writes bypass both
    # (PyTorch does not count them): call ``invalidate_native()`` after such an edit.
    def invalidate_native(self):
        self._native = None
        self._native_key = None
        self.__dict__.pop('_param_list', None)

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_native()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_native()
        return super().load_state_dict(*args, **kwargs)

    def _param_key(self):
        # (walking the module tree costs ~100 us per call; the module keeps its Parameter objects alive, so a
        #  data_ptr cannot be recycled by another tensor while it is part of the key)
        plist = self.__dict__.get('_param_list')
        if plist is None:
            plist = self.__dict__['_param_list'] = list(self.parameters())
        return tuple((p.data_ptr(), p._version) for p in plist)

    def native(self) -> _NativeDecoder:
        """Folds weight-norm on the host in fp32 (W = g v/||v||, exactly the tensor the
        reference's Linear multiplies with) and uploads; rebuilt when a parameter changes."""
        key = self._param_key()
        if self._native is None or key != self._native_key:
            weights, biases, lnw, lnb, concat, lnflag = [], [], [], [], [], []
            for l, _ in enumerate(self._layer_dims):
                lin = getattr(self, "lin" + str(l))
                if hasattr(lin, "weight_g"):
                    w = torch._weight_norm(lin.weight_v.detach().cpu().float(), lin.weight_g.detach().cpu().float(), 0)
                else:
                    w = lin.weight.detach().cpu().float()
                weights.append(w)
                biases.append(lin.bias.detach().cpu().float())
                has_ln = hasattr(self, "bn" + str(l))
                lnflag.append(1 if has_ln else 0)
                lnw.append(getattr(self, "bn" + str(l)).weight if has_ln else None)
                lnb.append(getattr(self, "bn" + str(l)).bias if has_ln else None)
                if l in self.latent_in:
                    concat.append(1)
                elif l != 0 and self.xyz_in_all:
                    concat.append(2)
                else:
                    concat.append(0)
            self._native = _NativeDecoder(self._layer_dims, concat, lnflag, self.latent_size, self.use_tanh,
                                          weights, biases, lnw, lnb)
            self._native_key = key
        return self._native

    # input: N x (L+3)
    def forward(self, input):
        if not input.is_cuda:
            raise _lib.SdfrError("sdflabel_b200.Decoder runs on a CUDA device only (no CPU path)")
        if self.training and ((self.dropout is not None and self.dropout_prob > 0) or self.latent_dropout):
            raise NotImplementedError("training-mode dropout is outside the render/refine path; call .eval()")
        x = _DecoderEval.apply(input, self.native(), self.mlp_impl)
        lat = input[:, :-3]
        if self.samples_per_scene:
            scale = self.scale_net(lat.reshape(-1, self.samples_per_scene, lat.size(1))[:, 0, :])
        else:
            scale = self.scale_net(lat[0])
        return x, scale
