"""Dev tool: the trace block of bench.py alone."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from sdflabel_b200.deepsdf.workspace import setup_dsdf
dev = torch.device("cuda")
sc = bench.load_scene()
dec, L = setup_dsdf(bench.PRIOR, precision=torch.float32); dec = dec.to(dev)
spec = json.load(open(os.path.splitext(bench.PRIOR)[0] + ".json"))
print(json.dumps(bench.run_trace(dec, sc, dev, bench.mlp_flops_per_point(spec), bench.peaks()), indent=1))
