"""Live cross-checks against the UNMODIFIED reference, where its tree is reachable (the build container; skipped on
the GPU box and anywhere else without /root/reference or SDFLABEL_REFERENCE).  Each check runs in a subprocess: the
reference needs its own sys.path entries and stubs for mpi4py / open3d / pyquaternion."""
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from oracle import ref_harness
from sdflabel_b200.pipelines import frames as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not reachable")


def _anno(i):
    return {'name': 'Car', 'bbox': np.array([10 + i, 20, 110 + i, 90]), 'alpha': 0.1 * i, 'rotation_y': 0.2 * i,
            'dimensions': np.array([1.5, 1.6, 3.9]), 'location': np.array([1.0 * i, 1.5, 10.0]), 'score': 1,
            'occluded': 0, 'truncated': 0.0}


def _label(i):
    return {'name': 'Car', 'bbox': np.array([10 + i, 20, 110 + i, 90]), 'location': np.array([1.0 * i, 1.4, 10.2]),
            'dimensions': [1.45, 1.62, 3.8], 'rotation_y': 0.21 * i, 'alpha': 0.11 * i, 'score': 1}


def test_label_dumps_through_the_reference_reader(tmp_path):
    """Our per-frame dumps read by the reference's OWN ``pipelines/evaluate_dump.py::evaluate`` (lines 20-46; the
    KITTI evaluator it hands the annotations to is replaced by a recorder) give what ``frames.load_autolabels``
    gives: same frames, same order, same arrays."""
    out = str(tmp_path / "labels")
    F.dump_frame_labels(out, 3, [_anno(0), _anno(1)], [_label(0), _label(1)])
    F.dump_frame_labels(out, 12, [_anno(2)], [])                 # every detection skipped: empty estimation
    F.dump_frame_labels(out, 7, [_anno(3), _anno(4), _anno(5)], [_label(3), _label(5)])
    captured = str(tmp_path / "captured.pkl")
    script = textwrap.dedent(f"""
        import configparser, os, pickle, sys, types
        import torch
        ref = {ref_harness.REF_ROOT!r}
        for p in (os.path.join(ref, "sdfrenderer"), ref):
            sys.path.insert(0, p)
        for name in ("open3d", "pyquaternion"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["pyquaternion"].Quaternion = object
        m = types.ModuleType("mpi4py")
        m.MPI = types.SimpleNamespace(COMM_WORLD=types.SimpleNamespace(Get_rank=lambda: 0))
        sys.modules["mpi4py"] = m
        torch.cuda.device_count = lambda: 1
        import pipelines.evaluate_dump as ED          # the reference's
        calls = []
        class Recorder:
            def __init__(self, *a, **k):
                pass
            def evaluate_detection_3d(self, gt, pred, classes, difficulties=None):
                calls.append((gt, pred))
                return "", {{}}
        ED.Detection3DEvaluator = Recorder
        cfgp = configparser.ConfigParser()
        cfgp.read_dict({{"output": {{"labels": {out!r}}}}})
        ED.evaluate(cfgp)
        pickle.dump(calls, open({captured!r}, "wb"))
    """)
    env = dict(os.environ, NUMBA_ENABLE_CUDASIM="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", script], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    calls = pickle.load(open(captured, "rb"))
    assert len(calls) == 2                                        # KITTI metrics, then the nuScenes metric
    gt_ref, pred_ref = calls[0]
    gt, pred = F.load_autolabels(out)
    assert [int(k) for k in gt] == [12, 3, 7] and len(gt_ref) == 3       # file-name order, like the reference
    for a, b in zip(gt_ref, gt.values()):
        assert set(a) == set(b)
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    for a, b in zip(pred_ref, pred.values()):
        assert set(a) == set(b)
        for key in a:
            assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    assert pred_ref[0]['name'] == [] and pred_ref[0]['location'].shape == (0, 3)
    assert pred_ref[2]['location'].shape == (2, 3)
