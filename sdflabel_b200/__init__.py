"""B200-native drop-in for the sdflabel SDF render / refine hot path.

Mirrors the Python object surface that pipelines/optimizer.py and
pipelines/refine_css.py of TRI-ML/sdflabel use:

    from sdflabel_b200.grid import Grid3D
    from sdflabel_b200.deepsdf.workspace import setup_dsdf
    from sdflabel_b200.renderer.rasterer import Rasterer
    from sdflabel_b200.pipelines.optimizer import Optimizer

All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
libsdfr.so (include/sdfr.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
