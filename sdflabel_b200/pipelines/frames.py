"""Frame sharding and the dump-time label exchange for multi-GPU refinement.

The reference processes frames in a serial Python loop (pipelines/refine_css.py:65)
and detections serially inside (refine_css.py:94); both are independent, and the
per-frame ``<idx>.pkl`` dump (refine_css.py:68-70,248) is its resume mechanism.
Here one process per GPU takes the frames ``i % world == rank``, batches all
detections of a frame into the same kernel launches (``BatchOptimizer``) and the
only collective is one all-gather of fixed-width label records at dump time so that
every rank (or rank 0) can run the evaluator (refine_css.py:253-263).  The payload is
O(100 B) per detection: latency-bound, NVLink bandwidth is irrelevant.

Record layout (float32): [frame, det, yaw, tx, ty, tz, scale, final_loss, latent(L)...].
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import numpy as np
import torch


def shard_frames(num_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame ownership (the natural unit: one dump file per frame)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, num_frames, world))


def record_width(latent_size: int) -> int:
    return 8 + latent_size


def make_record(frame: int, det: int, result: Dict, latent_size: int) -> np.ndarray:
    rec = np.zeros(record_width(latent_size), dtype=np.float32)
    hist = result.get('history')
    final = float(hist[-1, 2]) if hist is not None and len(hist) else float('nan')
    rec[0], rec[1] = frame, det
    rec[2] = result['yaw'][0]
    rec[3:6] = result['trans']
    rec[6] = result['scale'][0]
    rec[7] = final
    rec[8:] = result['latent']
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


def refine_frames(frames: Sequence[Sequence[Dict]], frame_ids: Iterable[int], dsdf, grid, weights, iters: int,
                  device='cuda') -> np.ndarray:
    """Refines the given frames (each a list of detection dicts, see BatchOptimizer) on this
    rank's GPU and returns their label records."""
    from .optimizer import BatchOptimizer
    L = dsdf.latent_size
    bo = BatchOptimizer(weights, device=device)
    recs = []
    for fid in frame_ids:
        dets = frames[fid]
        results = bo.optimize(iters, dets, dsdf, grid)
        for d, r in enumerate(results):
            recs.append(make_record(fid, d, r, L))
    return np.stack(recs) if recs else np.zeros((0, record_width(L)), dtype=np.float32)
