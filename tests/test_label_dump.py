"""Per-frame autolabel dumps: the reference's <idx>.pkl format (refine_css.py:241-248) and its reader
(evaluate_dump.py:20-46), restated here as the checker."""
import glob
import os
import pickle

import numpy as np

from sdflabel_b200.pipelines import frames as F


def _anno(i):
    return {'name': 'Car', 'bbox': np.array([10 + i, 20, 110 + i, 90]), 'alpha': 0.1 * i, 'rotation_y': 0.2 * i,
            'dimensions': np.array([1.5, 1.6, 3.9]), 'location': np.array([1.0 * i, 1.5, 10.0]), 'score': 1,
            'occluded': 0, 'truncated': 0.0}


def _label(i):
    return {'name': 'Car', 'bbox': np.array([10 + i, 20, 110 + i, 90]), 'location': np.array([1.0 * i, 1.4, 10.2]),
            'dimensions': [1.45, 1.62, 3.8], 'rotation_y': 0.21 * i, 'alpha': 0.11 * i, 'score': 1}


def _reference_reader(path_autolabels):
    """evaluate_dump.py:20-46, verbatim logic."""
    gt, pred = {}, {}
    for f in sorted(glob.glob(os.path.join(path_autolabels, '*.pkl'))):
        anno = pickle.load(open(f, "rb"))
        if 'skipped_frames' in f:
            continue
        frame_id = int(os.path.basename(f).split('.')[0])
        estimations = anno[1]
        if 'name' not in estimations:
            estimations['name'] = []
            estimations['location'] = np.zeros((0, 3))
            estimations['dimensions'] = np.zeros((0, 3))
            estimations['bbox'] = np.zeros((0, 4))
            estimations['rotation_y'] = np.zeros((0, ))
            estimations['alpha'] = np.zeros((0, ))
            estimations['score'] = np.zeros((0, ))
        gt[frame_id] = anno[0]
        pred[frame_id] = estimations
    return gt, pred


def test_dump_round_trip(tmp_path):
    out = str(tmp_path / "labels")
    assert not F.frame_done(out, 3)
    p = F.dump_frame_labels(out, 3, [_anno(0), _anno(1)], [_label(0), _label(1)])
    assert os.path.basename(p) == "3.pkl" and F.frame_done(out, 3)
    # a frame with annotations but no estimate (every detection skipped) still dumps; its estimation arrays are empty
    F.dump_frame_labels(out, 12, [_anno(2)], [])
    assert F.dump_frame_labels(out, 5, [], []) == '' and not F.frame_done(out, 5)     # refine_css.py:237
    gt, pred = _reference_reader(out)
    assert sorted(gt) == [3, 12]
    for key in F.NECESSARY_KEYS:
        assert isinstance(gt[3][key], np.ndarray) and isinstance(pred[3][key], np.ndarray)
        assert gt[3][key].shape[0] == 2 and pred[3][key].shape[0] == 2
    assert pred[3]['dimensions'].shape == (2, 3) and pred[3]['bbox'].shape == (2, 4) and pred[3]['name'] == ['Car', 'Car']
    assert np.allclose(pred[3]['location'][1], [1.0, 1.4, 10.2])
    assert pred[12]['name'] == [] and pred[12]['location'].shape == (0, 3)
    # our reader builds the same thing
    gt2, pred2 = F.load_autolabels(out)
    assert list(gt2) == [12, 3]                      # evaluate_dump.py sorts by file NAME ("12.pkl" < "3.pkl")
    for fid in gt:
        for key in F.NECESSARY_KEYS:
            assert np.array_equal(gt[fid][key], gt2[fid][key])
            assert np.array_equal(np.asarray(pred[fid][key]), np.asarray(pred2[fid][key]))
