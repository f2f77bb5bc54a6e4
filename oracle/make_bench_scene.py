"""Writes assets/bench_scene_256.npz: the synthetic cfg2 inputs bench.py feeds to both arms
(TEST INFRASTRUCTURE; run once in the build container: python -m oracle.make_bench_scene)."""
import os

import numpy as np

from . import prior as P
from . import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    prior = P.load_prior(os.path.join(ROOT, "assets", "deepsdf_synth.pt"))
    sc = scenes.make_scene(prior, size=256, density=40)
    out = {"K": sc["K"], "crop_size": np.asarray(sc["crop_size"]), "density": 40, "nocs_pred": sc["nocs_pred"],
           "lidar": sc["lidar"], "w2d": sc["weights"]["2d"], "w3d": sc["weights"]["3d"]}
    for k, v in sc["init"].items():
        out["init_" + k] = v
    for k, v in sc["gt"].items():
        out["gt_" + k] = v
    path = os.path.join(ROOT, "assets", "bench_scene_256.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
