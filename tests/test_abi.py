"""The C-ABI library loads on a CPU-only box and exports every symbol include/sdfr.h declares."""
import ctypes
import os
import re

import pytest

from sdflabel_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdfr.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdfr_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert os.path.isfile(_lib.LIB_PATH), "run `python -m sdflabel_b200._build` (or __graft_entry__.build())"


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdfr.h but not exported"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.sdfr_version() == 100
    assert isinstance(lib.sdfr_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sdflabel_b200.grid import Grid3D
    from sdflabel_b200.deepsdf.networks.deep_sdf_decoder_scale import Decoder
    with pytest.raises(_lib.SdfrError):
        Grid3D(8, device='cpu')
    dec = Decoder(3, [16, 16]).eval()
    with pytest.raises(_lib.SdfrError):
        dec(torch.zeros(4, 6))
    assert _lib.load().sdfr_caps() & 1 == 0


def test_product_never_imports_oracle():
    """The shipped package must not reference the test oracle."""
    pkg = os.path.join(ROOT, "sdflabel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)
