"""Frame sharding and the dump-time label exchange for multi-GPU refinement.

The reference processes frames in a serial Python loop (pipelines/refine_css.py:65)
and detections serially inside (refine_css.py:94); both are independent, and the
per-frame ``<idx>.pkl`` dump (refine_css.py:68-70,248) is its resume mechanism.
Here one process per GPU takes the frames ``i % world == rank``, batches all
detections of a frame into the same kernel launches (``BatchOptimizer``) and the
only collective is one all-gather of fixed-width label records at dump time so that
every rank (or rank 0) can run the evaluator (refine_css.py:253-263).  The payload is
O(100 B) per detection: latency-bound, NVLink bandwidth is irrelevant.

Record layout (float32): [frame, det, yaw, tx, ty, tz, scale, final_loss, latent(L)...,
dimensions(3), location(3), rotation_y, alpha] - the refined parameters and the KITTI label built from them.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import numpy as np
import torch


def shard_frames(num_frames: int, rank: int, world: int, detections_per_frame: Sequence[int] = None) -> List[int]:
    """Frames owned by ``rank`` (the natural unit: one dump file per frame, refine_css.py:65-70).

    Without ``detections_per_frame``: round robin.  With it (the annotation count of every frame is known before
    any GPU work starts): longest-processing-time-first - frames in decreasing detection count, each to the rank
    with the least work so far (ties: lowest rank) - because a frame costs time in proportion to its detections
    and frames carry 1 to 8 of them.  Deterministic, so every rank computes the same partition on its own."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    if detections_per_frame is None:
        return list(range(rank, num_frames, world))
    if len(detections_per_frame) != num_frames:
        raise ValueError("detections_per_frame must have one entry per frame")
    load = [0] * world
    mine: List[int] = []
    order = sorted(range(num_frames), key=lambda i: (-int(detections_per_frame[i]), i))
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += max(int(detections_per_frame[i]), 1)
        if r == rank:
            mine.append(i)
    return sorted(mine)


LABEL_FIELDS = 8     # dimensions(3), location(3), rotation_y, alpha


def record_width(latent_size: int) -> int:
    return 8 + latent_size + LABEL_FIELDS


def make_record(frame: int, det: int, result: Dict, latent_size: int) -> np.ndarray:
    rec = np.zeros(record_width(latent_size), dtype=np.float32)
    hist = result.get('history')
    final = float(hist[-1, 2]) if hist is not None and len(hist) else float('nan')
    rec[0], rec[1] = frame, det
    rec[2] = result['yaw'][0]
    rec[3:6] = result['trans']
    rec[6] = result['scale'][0]
    rec[7] = final
    rec[8:8 + latent_size] = result['latent']
    label = result.get('label')
    if label is not None:
        o = 8 + latent_size
        rec[o:o + 3] = label['dimensions']
        rec[o + 3:o + 6] = label['location']
        rec[o + 6], rec[o + 7] = label['rotation_y'], label['alpha']
    return rec


def gather_labels(local_records: np.ndarray, latent_size: int, device=None) -> np.ndarray:
    """All-gathers the (n_local, width) records of every rank and returns them sorted by
    (frame, det).  Works with any initialised torch.distributed backend (nccl on the GPUs,
    gloo in the CPU tests); without a process group it just sorts the local records."""
    width = record_width(latent_size)
    local = np.asarray(local_records, dtype=np.float32).reshape(-1, width)
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        allrec = local
    else:
        world = dist.get_world_size()
        dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
        count = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(count) for _ in range(world)]
        dist.all_gather(counts, count)
        cap = int(max(int(c.item()) for c in counts))
        buf = torch.zeros((cap, width), dtype=torch.float32, device=dev)
        if local.shape[0]:
            buf[:local.shape[0]] = torch.from_numpy(local).to(dev)
        bufs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(bufs, buf)
        allrec = np.concatenate([b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, counts)], axis=0)
    if allrec.shape[0] == 0:
        return allrec
    order = np.lexsort((allrec[:, 1], allrec[:, 0]))
    return allrec[order]


def checksum(records: np.ndarray) -> str:
    """Order-independent digest of a label set (T11: 1-GPU vs N-GPU runs must agree bit for bit)."""
    import hashlib
    rec = np.ascontiguousarray(records, dtype=np.float32)
    if rec.shape[0]:
        rec = rec[np.lexsort((rec[:, 1], rec[:, 0]))]
    return hashlib.sha256(rec.tobytes()).hexdigest()


def refine_frames(frames, frame_ids, dsdf, grid, weights, iters: int, device='cuda', path_autolabels=None,
                  max_batch=32, **kw) -> np.ndarray:
    """Refines the given frames (dicts with a ``detections`` list, see pipelines/refine_frames.py) on this
    rank's GPU - pose initialisation, refinement, labels, optional per-frame dumps - and returns their label
    records."""
    from .refine_frames import FrameRefiner, records_of
    fr = FrameRefiner(dsdf, grid, weights, iters, max_batch=max_batch, device=device, **kw)
    return records_of(fr.refine(frames, frame_ids, path_autolabels), dsdf.latent_size)


# ---------------------------------------------------------------------------------------------------
# Per-frame autolabel dumps in the reference's format
# ---------------------------------------------------------------------------------------------------
NECESSARY_KEYS = ('alpha', 'bbox', 'dimensions', 'location', 'rotation_y', 'score')   # refine_css.py:241


def frame_dump_path(path_autolabels: str, frame_idx: int) -> str:
    import os
    return os.path.join(path_autolabels, str(frame_idx) + '.pkl')


def frame_done(path_autolabels: str, frame_idx: int) -> bool:
    """The reference's resume rule: a frame whose dump exists is skipped (refine_css.py:68-70)."""
    import os
    return os.path.exists(frame_dump_path(path_autolabels, frame_idx))


def collect_labels(annos: Sequence[Dict], labels: Sequence[Dict]):
    """Lists of per-detection annotation / label dicts (``get_kitti_label`` output) -> the two dicts of
    arrays one frame dump holds (refine_css.py:90,232,241-245)."""
    from collections import defaultdict
    frame_annos, frame_estimations = defaultdict(list), defaultdict(list)
    for a in annos:
        for key, value in a.items():
            frame_annos[key].append(value)
    for l in labels:
        for key, value in l.items():
            frame_estimations[key].append(value)
    for key in NECESSARY_KEYS:
        frame_annos[key] = np.asarray(frame_annos[key])
        frame_estimations[key] = np.asarray(frame_estimations[key])
    return frame_annos, frame_estimations


def dump_frame_labels(path_autolabels: str, frame_idx: int, annos: Sequence[Dict], labels: Sequence[Dict]) -> str:
    """Writes ``<frame_idx>.pkl`` = pickle of ``[frame_annos, frame_estimations]``, the file
    ``pipelines/evaluate_dump.py:24-46`` parses and ``refine_css.py:68`` treats as "frame done"
    (refine_css.py:248).  Frames without annotations are not dumped (refine_css.py:237-238)."""
    import os
    import pickle
    if not annos:
        return ''
    os.makedirs(path_autolabels, exist_ok=True)
    frame_annos, frame_estimations = collect_labels(annos, labels)
    path = frame_dump_path(path_autolabels, frame_idx)
    tmp = path + '.tmp'
    with open(tmp, 'wb') as f:                       # atomic: a killed rank never leaves a half-written "done" marker
        pickle.dump([frame_annos, frame_estimations], f)
    os.replace(tmp, path)
    return path


def load_autolabels(path_autolabels: str):
    """(gt_annotations, pred_annotations) ordered by file name, exactly what evaluate_dump.py:20-46 builds
    for ``Detection3DEvaluator.evaluate_detection_3d``."""
    import glob
    import os
    import pickle
    from collections import OrderedDict
    gt, pred = OrderedDict(), OrderedDict()
    for f in sorted(glob.glob(os.path.join(path_autolabels, '*.pkl'))):
        if 'skipped_frames' in f:
            continue
        with open(f, 'rb') as fh:
            anno = pickle.load(fh)
        frame_id = int(os.path.basename(f).split('.')[0])
        estimations = anno[1]
        if 'name' not in estimations:
            estimations['name'] = []
            estimations['location'] = np.zeros((0, 3))
            estimations['dimensions'] = np.zeros((0, 3))
            estimations['bbox'] = np.zeros((0, 4))
            estimations['rotation_y'] = np.zeros((0,))
            estimations['alpha'] = np.zeros((0,))
            estimations['score'] = np.zeros((0,))
        gt[frame_id] = anno[0]
        pred[frame_id] = estimations
    return gt, pred
