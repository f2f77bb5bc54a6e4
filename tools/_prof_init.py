import cProfile, pstats, sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import synth_frames
from sdflabel_b200.deepsdf.workspace import setup_dsdf
from sdflabel_b200.grid import Grid3D
from sdflabel_b200.pipelines.refine_frames import FrameRefiner
dev = torch.device("cuda")
dec, L = setup_dsdf(os.path.join(ROOT, "assets", "deepsdf_synth.pt"), precision=torch.float32)
dec = dec.to(dev)
grid = Grid3D(40, device=dev)
frames = synth_frames.make_frames(96, seed=0)
fr = FrameRefiner(dec, grid, {"2d": 0.3, "3d": 0.5}, iters=60, max_batch=32)
fr.refine(synth_frames.make_frames(2, seed=99), [0, 1])
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
fr.refine(frames, list(range(len(frames))))
torch.cuda.synchronize()
pr.disable()
print("wall", time.perf_counter() - t0, "detections", sum(len(f["detections"]) for f in frames), fr.timing)
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
