"""ctypes binding of libsdfr.so (the C ABI declared in include/sdfr.h).

The product path has no CPU implementation: if the library is missing or no CUDA
device is visible, every entry point raises instead of silently doing something else.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# SDFR_LIB points the binding at another build of the same ABI (a profiling build, a kernel variant under test)
LIB_PATH = os.environ.get("SDFR_LIB") or os.path.join(HERE, "libsdfr.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int32)
vp = C.c_void_p

MLP_AUTO, MLP_FFMA, MLP_TCGEN05, MLP_TCGEN05_COARSE = 0, 1, 2, 3
ROT_DCM, ROT_QUAT = 0, 1
PRIM_DISC, PRIM_CIRCLE, PRIM_CIRCLE_OPT = 0, 1, 2


class SdfrError(RuntimeError):
    pass


class DecoderSpec(C.Structure):
    _fields_ = [
        ("latent_size", C.c_int32), ("num_layers", C.c_int32),
        ("in_dims", c_int_p), ("out_dims", c_int_p), ("concat", c_int_p), ("layer_norm", c_int_p),
        ("use_tanh", C.c_int32),
    ]


class RasterCfg(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("kinv", C.c_float * 9), ("k", C.c_float * 9),
        ("rot", C.c_int32), ("output_nocs", C.c_int32), ("primitive", C.c_int32), ("bg_dev", C.c_void_p),
    ]


class RefineCfg(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("density", C.c_int32), ("max_width", C.c_int32), ("max_height", C.c_int32),
        ("max_lidar", C.c_int32), ("max_iters", C.c_int32), ("weight_2d", C.c_float), ("weight_3d", C.c_float),
        ("mlp_impl", C.c_int32), ("latent_lipschitz", C.c_float),
    ]


# name -> (restype, argtypes); every symbol include/sdfr.h declares
SIGNATURES = {
    "sdfr_version": (C.c_int, []),
    "sdfr_last_error": (C.c_char_p, []),
    "sdfr_caps": (C.c_int, []),
    "sdfr_launch_count": (C.c_int64, []),
    "sdfr_decoder_create": (C.c_int, [C.POINTER(DecoderSpec), C.POINTER(c_float_p), C.POINTER(c_float_p),
                                      C.POINTER(c_float_p), C.POINTER(c_float_p), C.POINTER(vp)]),
    "sdfr_decoder_destroy": (None, [vp]),
    "sdfr_decoder_tcgen05_ok": (C.c_int, [vp]),
    "sdfr_decoder_check": (C.c_int, [vp]),
    "sdfr_decoder_eval": (C.c_int, [vp, vp, C.c_int64, vp, vp, C.c_int, vp]),
    "sdfr_decoder_eval_lattice": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "sdfr_lattice_points": (C.c_int, [C.c_int, vp, vp]),
    "sdfr_surface_extract": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_float, vp, vp, vp, vp,
                                       vp, vp]),
    "sdfr_splat_workspace_bytes": (C.c_int64, [C.POINTER(RasterCfg), C.c_int64]),
    "sdfr_splat_forward": (C.c_int, [C.POINTER(RasterCfg), vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp,
                                     vp, vp, vp, vp]),
    "sdfr_splat_backward": (C.c_int, [C.POINTER(RasterCfg), vp, vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp,
                                      vp, vp, vp, vp]),
    "sdfr_trace_workspace_bytes": (C.c_int64, [C.POINTER(RasterCfg), vp]),
    "sdfr_trace_forward": (C.c_int, [vp, C.POINTER(RasterCfg), vp, c_float_p, C.c_int, C.c_float, vp, vp, vp, vp, vp, vp,
                                     vp, C.c_float, C.c_int, vp]),
    "sdfr_trace_cache_bytes": (C.c_int64, []),
    "sdfr_trace_cache_update": (C.c_int, [vp, vp, vp, C.c_float, vp]),
    "sdfr_trace_set_stats": (None, [C.c_int]),
    "sdfr_trace_get_stats": (None, [C.POINTER(C.c_int64)]),
    "sdfr_trace_backward": (C.c_int, [vp, C.POINTER(RasterCfg), c_float_p, C.c_float, vp, vp, vp, vp, vp, vp]),
    "sdfr_loss3d": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_double, vp, vp, vp, vp]),
    "sdfr_loss2d": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, vp]),
    "sdfr_refine_create": (C.c_int, [vp, C.POINTER(RefineCfg), C.POINTER(vp)]),
    "sdfr_refine_destroy": (None, [vp]),
    "sdfr_refine_set_detection": (C.c_int, [vp, C.c_int, c_float_p, c_float_p, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                            vp, C.c_int, c_float_p, c_float_p, c_float_p, c_float_p, vp]),
    "sdfr_refine_set_active": (C.c_int, [vp, C.c_int]),
    "sdfr_refine_import": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp]),
    "sdfr_refine_set_optimizer_state": (C.c_int, [vp, C.c_int, c_float_p, c_float_p, C.c_int, vp]),
    "sdfr_refine_get_optimizer_state": (C.c_int, [vp, C.c_int, c_float_p, c_float_p, C.POINTER(C.c_int)]),
    "sdfr_refine_run": (C.c_int, [vp, C.c_int, vp]),
    "sdfr_refine_get_batch": (C.c_int, [vp, c_float_p, c_float_p, C.POINTER(C.c_int), vp]),
    "sdfr_refine_preselect_error": (C.c_int, [vp, c_float_p, vp]),
    "sdfr_refine_label_extents": (C.c_int, [vp, c_float_p, vp]),
    "sdfr_refine_set_latent": (C.c_int, [vp, C.c_int, c_float_p, vp]),
    "sdfr_refine_profile": (C.c_int, [vp, C.c_int, c_float_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int32), vp]),
    "sdfr_refine_stage_name": (C.c_char_p, [C.c_int]),
    "sdfr_refine_lattice_rows": (C.c_int, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, vp]),
    "sdfr_refine_get": (C.c_int, [vp, C.c_int, c_float_p, c_float_p, C.POINTER(C.c_int), vp]),
    "sdfr_refine_export": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp]),
    "sdfr_refine_optimize": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int,
                                       vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]),
    "sdfr_refine_view": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]),
    "sdfr_refine_copy_view": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int64, vp]),
    "sdfr_rotate_iou": (C.c_int, [vp, C.c_int64, vp, C.c_int64, C.c_int, vp, vp]),
    "sdfr_np_choice4": (C.c_int, [vp, C.POINTER(C.c_int32), C.c_int64, C.c_int32, vp]),
    "sdfr_nn_query": (C.c_int, [vp, C.c_int64, vp, C.c_int64, vp, vp, vp]),
    "sdfr_ransac_score": (C.c_int, [vp, vp, C.c_int64, vp, vp, C.c_int64, vp, C.c_int, C.c_double, C.c_float, vp, vp,
                                    vp]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Loads libsdfr.so (raises SdfrError when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise SdfrError(
                f"{LIB_PATH} is missing: build it with `python -m sdflabel_b200._build` "
                "(there is no CPU fallback for the render/refine path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sdfr_last_error().decode("utf-8", "replace")
        raise SdfrError(f"libsdfr error {rc}: {msg}")


def require_cuda() -> None:
    lib = load()
    if not (lib.sdfr_caps() & 1):
        raise SdfrError("no CUDA device visible: sdflabel_b200 has no CPU path")


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device.  The raw accessor is used when this torch has it
    (0.3 us instead of the 5 - 15 us of building a torch.cuda.Stream object: the refine loop asks six times per call)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    """data_ptr of a tensor, or 0 for None."""
    return 0 if t is None else t.data_ptr()


def fptr(arr):
    """float* view of a contiguous float32 numpy array (host)."""
    return arr.ctypes.data_as(c_float_p)
