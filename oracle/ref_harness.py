"""Imports the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Only usable in the build container: the GPU box has no /root/reference, so
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this module.
It exists for ``oracle/make_golden.py`` (which writes tests/golden/*.npz) and for
the optional live cross-check ``tests/test_oracle_live_reference.py``.

Recipe verified in SURVEY.md Appendix B: both the repo root and ``sdfrenderer/``
must be on sys.path (README.md:21-24, workspace.py:172) and the two viz-only
imports ``open3d`` / ``pyquaternion`` are stubbed (optimizer.py:1, refinement.py:4,6).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("SDFLABEL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "pipelines", "optimizer.py"))


_loaded = None


def load():
    """Returns a namespace with the reference's Grid3D, Rasterer, Decoder,
    setup_dsdf, Optimizer and rot_from_yaw."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for p in (os.path.join(REF_ROOT, "sdfrenderer"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("open3d", "pyquaternion"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["pyquaternion"], "Quaternion"):
        sys.modules["pyquaternion"].Quaternion = object
    import grid as ref_grid                                   # noqa: E402
    from renderer.rasterer import Rasterer                    # noqa: E402
    import deepsdf.workspace as ws                            # noqa: E402
    from deepsdf.networks.deep_sdf_decoder_scale import Decoder  # noqa: E402
    from pipelines.optimizer import Optimizer                 # noqa: E402
    import utils.refinement as rtools                         # noqa: E402
    ns = types.SimpleNamespace(
        grid_module=ref_grid, Grid3D=ref_grid.Grid3D, Rasterer=Rasterer, Decoder=Decoder,
        setup_dsdf=ws.setup_dsdf, Optimizer=Optimizer, rot_from_yaw=rtools.rot_from_yaw,
    )
    _loaded = ns
    return ns
