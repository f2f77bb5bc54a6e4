"""N > 1 host logic on CPU: frame sharding + the dump-time label all-gather over gloo (world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from sdflabel_b200.pipelines import frames as F

L = 3
NUM_FRAMES = 7


def _fake_records(frame_ids):
    recs = []
    for f in frame_ids:
        rng = np.random.RandomState(100 + f)
        for d in range(f % 3 + 1):
            r = {'yaw': rng.rand(1).astype(np.float32), 'trans': rng.rand(3).astype(np.float32),
                 'scale': rng.rand(1).astype(np.float32), 'latent': rng.rand(L).astype(np.float32),
                 'history': rng.rand(4, 4).astype(np.float32)}
            recs.append(F.make_record(f, d, r, L))
    return np.stack(recs) if recs else np.zeros((0, F.record_width(L)), dtype=np.float32)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = F.shard_frames(NUM_FRAMES, rank, world, [f % 3 + 1 for f in range(NUM_FRAMES)])
    allrec = F.gather_labels(_fake_records(mine), L, device='cpu')
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), allrec)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_covers_all_frames_once():
    for world in (1, 2, 3, 8):
        seen = sorted(sum((F.shard_frames(NUM_FRAMES, r, world) for r in range(world)), []))
        assert seen == list(range(NUM_FRAMES))


def test_shard_by_detection_count_is_balanced_and_deterministic():
    """Longest-processing-time-first over the per-frame detection counts: every frame exactly once, the ranks'
    detection totals within one frame's worth of each other, the same partition whoever computes it."""
    rng = np.random.RandomState(0)
    counts = rng.randint(1, 9, size=512).tolist()
    for world in (1, 2, 4, 8):
        parts = [F.shard_frames(512, r, world, counts) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(512))
        loads = [sum(counts[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= 8, (world, loads)
        assert parts == [F.shard_frames(512, r, world, list(counts)) for r in range(world)]
        rr = [sum(counts[i] for i in F.shard_frames(512, r, world)) for r in range(world)]
        assert max(loads) <= max(rr)
    # frames without detections still belong to somebody
    parts = [F.shard_frames(5, r, 2, [0, 0, 3, 0, 1]) for r in range(2)]
    assert sorted(sum(parts, [])) == [0, 1, 2, 3, 4]


def test_record_layout_carries_the_label():
    r = {'yaw': np.float32([0.5]), 'trans': np.float32([1, 2, 3]), 'scale': np.float32([2]), 'latent': np.float32([.1, .2, .3]),
         'history': np.float32([[0, 0, 0.25, 0]]),
         'label': {'dimensions': [1.5, 1.6, 4.0], 'location': np.array([1., 2., 30.]), 'rotation_y': -1.2, 'alpha': -1.1}}
    rec = F.make_record(7, 2, r, 3)
    assert rec.shape == (F.record_width(3),) == (19,)
    assert rec[0] == 7 and rec[1] == 2 and rec[7] == 0.25 and np.allclose(rec[11:14], [1.5, 1.6, 4.0])
    assert np.allclose(rec[14:17], [1, 2, 30]) and np.allclose(rec[17:], [-1.2, -1.1])


def test_gather_world2_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = F.gather_labels(_fake_records(range(NUM_FRAMES)), L)
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npy")
        assert got.shape == single.shape
        assert np.array_equal(got, single)           # bit exact per record, same order
        assert F.checksum(got) == F.checksum(single)


def test_gather_handles_empty_rank(tmp_path):
    # more ranks than frames: some ranks contribute nothing
    global NUM_FRAMES
    recs = F.gather_labels(_fake_records([]), L)
    assert recs.shape == (0, F.record_width(L))


# ---- sharded per-frame dumps + the reference's resume rule (refine_css.py:68-70, 241-248) ----------------
def _dump_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for f in F.shard_frames(NUM_FRAMES, rank, world):
        if F.frame_done(out_dir, f):                 # resume: frame 2 was dumped by an earlier (interrupted) run
            continue
        n = f % 3 + 1
        annos = [{'name': 'Car', 'bbox': np.array([f, d, f + 50, d + 40]), 'alpha': 0.1 * d, 'rotation_y': 0.2 * d,
                  'dimensions': np.array([1.5, 1.6, 3.9]), 'location': np.array([1.0 * f, 1.5, 10.0 + d]), 'score': 1}
                 for d in range(n)]
        labels = [{'name': 'Car', 'bbox': a['bbox'], 'location': a['location'] + 0.1, 'dimensions': [1.4, 1.6, 3.8],
                   'rotation_y': a['rotation_y'] + 0.01, 'alpha': a['alpha'] + 0.01, 'score': 1} for a in annos]
        F.dump_frame_labels(out_dir, f, annos, labels)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_dumps_cover_every_frame_and_resume(tmp_path):
    out = str(tmp_path / "labels")
    marker = F.dump_frame_labels(out, 2, [{'name': 'Car', 'bbox': np.zeros(4), 'alpha': 0., 'rotation_y': 0.,
                                          'dimensions': np.ones(3), 'location': np.zeros(3), 'score': 1}], [])
    stamp = os.path.getmtime(marker)
    world = 2
    mp.spawn(_dump_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    gt, pred = F.load_autolabels(out)
    assert sorted(gt) == list(range(NUM_FRAMES))                     # every frame exactly once, whoever owned it
    assert os.path.getmtime(marker) == stamp and pred[2]['name'] == []      # the finished frame was not redone
    for f in range(NUM_FRAMES):
        if f != 2:
            assert pred[f]['location'].shape == (f % 3 + 1, 3) and gt[f]['bbox'].shape == (f % 3 + 1, 4)
    assert not [p for p in os.listdir(out) if p.endswith('.tmp')]     # atomic writes leave nothing behind


def test_batch_plan_covers_every_detection_once_in_order():
    """refine_frames.plan_batches: whole frames per batch up to the limit, oversized frames sliced, the first batch a
    quarter of the limit (pipeline fill), every detection exactly once and in (frame, detection) order."""
    from sdflabel_b200.pipelines.refine_frames import plan_batches
    rng = np.random.RandomState(3)
    for max_batch in (1, 5, 32):
        counts = rng.randint(0, 9, size=40).tolist() + [70, 0, 33, 1]
        frames = [{'detections': [{'id': (f, d)} for d in range(c)]} for f, c in enumerate(counts)]
        todo = [i for i in range(len(frames)) if i % 7 != 3]
        batches = plan_batches(frames, todo, max_batch)
        flat = [(fid, di) for b in batches for fid, di, det in b]
        assert flat == [(fid, di) for fid in todo for di in range(counts[fid])]
        assert all(det is frames[fid]['detections'][di] for b in batches for fid, di, det in b)
        assert all(0 < len(b) <= max_batch for b in batches)
        assert len(batches[0]) <= max(1, min(max_batch, max(4, max_batch // 4)))
        # a frame that fits a batch is never split across two
        for b in batches:
            for fid in {f for f, _, _ in b}:
                if counts[fid] <= max(1, min(max_batch, max(4, max_batch // 4))):
                    assert sum(1 for f, _, _ in b if f == fid) == counts[fid]
    assert plan_batches([], [], 8) == [] and plan_batches([{'detections': []}], [0], 8) == []
