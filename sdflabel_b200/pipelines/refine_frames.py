"""The frame / detection loop of the reference's ``refine_css`` on one GPU.

Mirror of pipelines/refine_css.py:65-250 from the point where the CSS network has spoken: per
detection the initial pose (model cloud from the predicted latent -> ``PoseEstimator`` ->
yaw / translation / scale, refine_css.py:143-199), the refinement (``Optimizer.optimize``,
203-226), the KITTI label (``get_kitti_label``, 229-231) and the per-frame ``<idx>.pkl`` dump with
its resume rule (68-70, 241-250).  What is NOT here is the data layer above it (KITTI loading,
crops, the CSS network): a *detection* arrives as a dict of what those stages produce

    K (3,3) crop intrinsics     crop_size [H, W]      nocs_pred (3,h,w)      lidar (N,3)
    latent_pred (L,)            scene_pts / scene_cls (S,3): NOCS-coloured scene cloud
    orig_cam (3,3), bbox [l,t,r,b], anno (dict; optional)

and a *frame* as ``{"detections": [...], "world_to_cam": 4x4}``.

The reference walks frames and detections serially.  Here every stage is batched over all
detections of as many consecutive frames as fit ``max_batch`` engine slots:

  * one lattice evaluation + surface extraction for all predicted latents (the model clouds),
  * the pose RANSAC per detection (its kernels run on a side stream, its host part overlaps the
    previous batch's refinement: the engine's 60 iterations are enqueued asynchronously),
  * one refinement of the whole batch, one read-back, one label-extent pass.

Every result is independent of how detections are grouped (bit for bit: tests/test_gpu_frames.py),
which is what lets the multi-GPU driver shard frames freely.
"""
from __future__ import annotations

import math
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import frames as F
from .optimizer import BatchOptimizer, _Engine, _pow2_ceil
from .. import _lib
from ..utils import refinement as rtools
from ..utils.pose import PoseEstimator


def project(K, p3d):
    """Pinhole projection without distortion (the reference's cv2.projectPoints call with zero rvec / tvec,
    utils/refinement.py:470-472)."""
    p = np.asarray(p3d, dtype=np.float64) @ np.asarray(K, dtype=np.float64).T
    return (p[:, :2] / p[:, 2:3]).astype(np.float32)


def compute_iou(box_a, box_b):
    """utils/refinement.py:168-198 (inclusive pixel boxes)."""
    xa, ya = max(box_a[0], box_b[0]), max(box_a[1], box_b[1])
    xb, yb = min(box_a[2], box_b[2]), min(box_a[3], box_b[3])
    inter = max(0, xb - xa + 1) * max(0, yb - ya + 1)
    area_a = (box_a[2] - box_a[0] + 1) * (box_a[3] - box_a[1] + 1)
    area_b = (box_b[2] - box_b[0] + 1) * (box_b[3] - box_b[1] + 1)
    return inter / float(area_a + area_b - inter)


def initial_params(det: Dict, model_pts: torch.Tensor, estimator: PoseEstimator, rng=None) -> Optional[Dict]:
    """refine_css.py:150-199: RANSAC pose from the NOCS correspondences, rotation constrained to the
    azimuth, height re-estimated when the projected model misses the 2D box.  ``model_pts`` are the
    isosurface points of the predicted latent (device tensor, object frame); ``rng`` as in
    ``PoseEstimator.estimate``."""
    nocs_dsdf = (model_pts + 1) / 2                                     # grid.py:67
    init_pose = estimator.estimate(model_pts, nocs_dsdf, det['scene_pts'], det['scene_cls'], None, None, rng=rng)
    if init_pose is None:
        return None                                                     # 'NO RANSAC POSE FOUND!!!' (171-173)
    scale, rot, tra = init_pose['scale'], np.array(init_pose['rot'], dtype=np.float64), np.array(init_pose['tra'])
    rot[:, 1] = [0, 1, 0]
    rot[1, :] = [0, 1, 0]
    yaw = rtools.roty_in_bev(rot @ np.diag([-1, 1, 1])) + math.pi / 2   # KITTI roty starts at canonical pi/2
    if det.get('orig_cam') is not None and det.get('bbox') is not None:
        world_points = (rot @ (model_pts.detach().cpu().numpy() * scale).T).T + tra
        proj = project(det['orig_cam'], world_points)
        box = [proj[:, 0].min(), proj[:, 1].min(), proj[:, 0].max(), proj[:, 1].max()]
        if compute_iou(list(det['bbox']), box) < 0.7:                   # 'Restimating height' (183-186)
            ymin, ymax = world_points[:, 1].min(), world_points[:, 1].max()
            tra[1] = np.asarray(det['scene_pts'])[:, 1].min() + (ymax - ymin) / 2
    return {'yaw': np.array([yaw], dtype=np.float32), 'trans': (tra / scale).astype(np.float32),
            'scale': np.array([scale], dtype=np.float32),
            'latent': np.asarray(det['latent_pred'], dtype=np.float32).copy()}


def plan_batches(frames: Sequence[Dict], todo: Sequence[int], max_batch: int) -> List[List]:
    """Batches of whole frames, at most ``max_batch`` detections each (a larger frame gets batches of its own slices),
    as lists of (frame_id, detection_index, detection).  The FIRST batch is a quarter of that: nothing overlaps its
    initialisation (the GPU waits for it), and with the frames of a run sharded over many ranks that pipeline fill is
    a visible share of a rank's few batches.  Results do not depend on the grouping."""
    batches, cur = [], []
    for fid in todo:
        dets = frames[fid]['detections']
        items = [(fid, di, d) for di, d in enumerate(dets)]
        limit = max_batch if batches else max(1, min(max_batch, max(4, max_batch // 4)))
        while len(items) > limit:
            if cur:
                batches.append(cur)
                cur = []
            batches.append(items[:limit])
            items = items[limit:]
            limit = max_batch
        if len(cur) + len(items) > limit:
            batches.append(cur)
            cur = []
        cur = cur + items
    if cur:
        batches.append(cur)
    return batches


class FrameRefiner:
    """Refines frames on this process's GPU; see the module docstring."""

    def __init__(self, dsdf, grid, weights, iters, max_batch=32, max_crop=(96, 96), max_lidar=1024,
                 pose_estimator='kabsch', init_scale=2.0, device='cuda', seed=0, init_threads=None):
        self.dsdf, self.grid, self.weights, self.iters = dsdf, grid, weights, int(iters)
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:      # pin the index now: the current device is per THREAD, the workers start on 0
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.max_batch = int(max_batch)
        self.estimator = PoseEstimator(pose_estimator, init_scale)
        self.seed = int(seed)
        self.bo = BatchOptimizer(weights, device=self.device)
        with torch.cuda.device(self.device):
            self.engine = self.bo.reserve(dsdf, grid, self.max_batch, max_crop, max_lidar, self.iters)
            # a second, small engine evaluates the model clouds of the NEXT batch while this one refines
            self.init_engine = _Engine(dsdf.native(), _pow2_ceil(self.max_batch, 1), int(grid.density), 32, 32, 1, 1,
                                       float(weights['2d']), float(weights['3d']),
                                       getattr(dsdf, 'mlp_impl', _lib.MLP_AUTO))
            # (the tensor-core decoder keeps its scratch per engine; the CUDA-core kernel's LayerNorm scratch is
            #  per decoder, so with that kernel the two engines must not overlap: same stream)
            # High priority: the initialisation kernels are short and the host waits for each of them, the refinement
            # is a long asynchronous train of launches; pending blocks of the side streams go first.
            self.concurrent_init = bool(dsdf.native().tcgen05)
            self.init_stream = torch.cuda.Stream(device=self.device, priority=-1) if self.concurrent_init else \
                torch.cuda.current_stream(self.device)
        # the host part of the pose initialisation (sample draw, 4-point fits, read-backs) runs on a few threads:
        # its C / LAPACK / CUDA calls release the GIL; every detection draws from a generator of its own
        if init_threads is None:
            local_ranks = max(1, int(os.environ.get('LOCAL_WORLD_SIZE', '1')))
            init_threads = max(1, min(6, (os.cpu_count() or 2) // local_ranks - 1))
        self.init_threads = int(init_threads)
        self.pool = ThreadPoolExecutor(max_workers=self.init_threads) if self.init_threads > 1 else None
        # every worker thread initialises its detections on a stream of its own: its read-backs wait for its own
        # kernels only
        self._tls = threading.local()
        self._events = []           # (start, end) CUDA events around the refinement of each batch
        self.timing = {'init_s': 0.0, 'refine_wait_s': 0.0, 'label_s': 0.0, 'refine_gpu_s': 0.0, 'detections': 0,
                       'batches': 0, 'no_pose': 0}

    # ---- stage 1: model clouds + pose RANSAC of a batch (side stream) ---------------------------------
    def _prepare(self, items):
        """items: [(frame_id, det_index, det)] -> the same list with 'params' (or None) per entry."""
        import time
        t0 = time.perf_counter()
        eng = self.init_engine
        out = []
        with torch.cuda.device(self.device), torch.cuda.stream(self.init_stream):
            for s in range(0, len(items), eng.cfg.batch):
                chunk = items[s:s + eng.cfg.batch]
                lat = np.stack([np.asarray(d['latent_pred'], dtype=np.float32) for _, _, d in chunk])
                clouds = eng.surface_clouds(lat)
                self._clouds_ready = torch.cuda.Event()
                self._clouds_ready.record(self.init_stream)
                work = list(zip(chunk, clouds))
                params = list(self.pool.map(self._initial, work)) if self.pool else [self._initial(w) for w in work]
                out.extend((fid, di, det, p) for ((fid, di, det), _), p in zip(work, params))
        self.timing['init_s'] += time.perf_counter() - t0
        return out

    def _initial(self, item):
        (fid, di, det), cloud = item
        if det.get('params') is not None:
            return det['params']
        # the reference consumes numpy's global RNG in file order; a generator seeded per detection keeps a
        # detection's hypotheses independent of which rank / batch / thread it lands in
        rng = np.random.RandomState((self.seed * 1000003 + fid * 131 + di) % (2 ** 32))
        stream = self.init_stream
        if self.pool is not None and self.concurrent_init:
            stream = getattr(self._tls, 'stream', None)
            if stream is None:
                stream = self._tls.stream = torch.cuda.Stream(device=self.device, priority=-1)
            stream.wait_event(self._clouds_ready)
        with torch.cuda.device(self.device), torch.cuda.stream(stream):
            return initial_params(det, cloud, self.estimator, rng)

    # ---- stage 2: refinement of a prepared batch (main stream, asynchronous) ---------------------------
    def _launch(self, prepared):
        live = [(fid, di, det, p) for fid, di, det, p in prepared if p is not None]
        self.timing['no_pose'] += len(prepared) - len(live)
        if not live:
            return live
        eng = self.engine
        with torch.cuda.device(self.device):
            if eng.active != len(live):
                eng.set_active(len(live))
            for b, (fid, di, det, p) in enumerate(live):
                eng.set_detection(b, det['K'], int(det['crop_size'][1]), int(det['crop_size'][0]), det['nocs_pred'],
                                  det['lidar'], p['yaw'], p['trans'], p['scale'], p['latent'])
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            eng.run(self.iters)
            ev[1].record()
            self._events.append(ev)
        return live

    # ---- stage 3: read-back, labels ----------------------------------------------------------------------
    def _finish(self, live, frames):
        import time
        if not live:
            return []
        eng = self.engine
        t0 = time.perf_counter()
        with torch.cuda.device(self.device):
            out, hists = eng.get_batch()
            self.timing['refine_wait_s'] += time.perf_counter() - t0
            for a, b in self._events:                       # device time from the first to the last launch of the batch
                self.timing['refine_gpu_s'] += a.elapsed_time(b) * 1e-3
            self._events = []
            t1 = time.perf_counter()
            ext = eng.label_extents()
        results = []
        for b, (fid, di, det, p) in enumerate(live):
            res = {'yaw': out[b, 0:1].copy(), 'trans': out[b, 1:4].copy(), 'scale': out[b, 4:5].copy(),
                   'latent': out[b, 5:].copy(), 'history': hists[b].copy()}
            label, cam_T = rtools.kitti_label_from_extents(
                ext[b, 0:3], ext[b, 3:6], res['latent'], res['scale'], res['trans'], res['yaw'],
                frames[fid].get('world_to_cam', np.eye(4)), det.get('bbox', [0, 0, 0, 0]))
            res['label'] = label
            results.append((fid, di, res))
        self.timing['label_s'] += time.perf_counter() - t1
        self.timing['detections'] += len(live)
        self.timing['batches'] += 1
        return results

    def refine(self, frames: Sequence[Dict], frame_ids: Iterable[int], path_autolabels: Optional[str] = None):
        """Refines ``frames[i]`` for i in ``frame_ids`` and returns {frame_id: [result per detection]} (a result
        is None where no initial pose was found).  With ``path_autolabels`` every finished frame is dumped as
        ``<idx>.pkl`` and frames whose dump exists are skipped (the reference's resume rule)."""
        todo = [i for i in frame_ids if not (path_autolabels and F.frame_done(path_autolabels, i))]
        batches = plan_batches(frames, todo, self.max_batch)
        done: Dict[int, List] = {fid: [None] * len(frames[fid]['detections']) for fid in todo}
        remaining = {fid: len(frames[fid]['detections']) for fid in todo}
        for fid in todo:
            if remaining[fid] == 0:
                self._frame_complete(fid, frames, done, path_autolabels)

        prepared = self._prepare(batches[0]) if batches else None
        for k in range(len(batches)):
            self.init_stream.synchronize()
            live = self._launch(prepared)                   # asynchronous: the GPU refines while ...
            prepared = self._prepare(batches[k + 1]) if k + 1 < len(batches) else None   # ... the next batch is initialised
            for fid, di, res in self._finish(live, frames):
                done[fid][di] = res
            for fid, di, det in batches[k]:
                remaining[fid] -= 1
                if remaining[fid] == 0:
                    self._frame_complete(fid, frames, done, path_autolabels)
        return done

    def _frame_complete(self, fid, frames, done, path_autolabels):
        if not path_autolabels:
            return
        annos, labels = [], []
        for di, det in enumerate(frames[fid]['detections']):
            annos.append(det.get('anno', {'bbox': det.get('bbox', [0, 0, 0, 0])}))
            if done[fid][di] is not None:
                labels.append(done[fid][di]['label'])
        F.dump_frame_labels(path_autolabels, fid, annos, labels)


def records_of(done: Dict[int, List], latent_size: int) -> np.ndarray:
    """Fixed-width label records (frames.make_record) of a ``FrameRefiner.refine`` result."""
    recs = [F.make_record(fid, di, r, latent_size) for fid in sorted(done) for di, r in enumerate(done[fid])
            if r is not None]
    return np.stack(recs) if recs else np.zeros((0, F.record_width(latent_size)), dtype=np.float32)
