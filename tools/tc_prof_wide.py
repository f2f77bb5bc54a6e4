"""In-kernel phase timers of mlp_tc_coarse_wide_kernel (build with SDFR_NVCC_FLAGS=-DSDFR_TC_PROFILE)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdflabel_b200 import _lib
from sdflabel_b200.deepsdf.workspace import setup_dsdf
lib = _lib.load()
dev = torch.device("cuda")
stock, L = setup_dsdf(os.path.join(ROOT, "assets", "deepsdf_synth.pt"), precision=torch.float32)
stock = stock.to(dev)
lat = torch.nn.functional.normalize(torch.tensor([[0.5, 0.7, 0.5]]), dim=1).to(dev)
sdf = torch.empty(64000, device=dev)
prof = lib.sdfr_debug_tc_prof
prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
for it in range(3):
    prof(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.sdfr_decoder_eval_lattice(stock.native().handle, lat.data_ptr(), 1, 40, sdf.data_ptr(), 0, _lib.MLP_TCGEN05_COARSE, _lib.stream_ptr()))
    b.record(); torch.cuda.synchronize()
    prof(buf, 0)
    names = {8: "producer wait empty", 9: "mma wait act(epilogue)", 10: "mma wait full(weights)", 11: "epi wait acc early(mb<hold)", 12: "epi wait acc rest", 13: "issuer loop total"}
    print(f"run {it}: {a.elapsed_time(b):.3f} ms; CTA0 cycles:", {n: int(buf[i]) for i, n in names.items()}, flush=True)
