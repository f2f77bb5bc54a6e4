"""In-kernel phase timers of mlp_tc_kernel (library must be built with SDFR_NVCC_FLAGS=-DSDFR_TC_PROFILE)."""
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
sdf = torch.empty(64000, device=dev); din = torch.empty(64000, 6, device=dev)
prof = lib.sdfr_debug_tc_prof
prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
for it in range(3):
    prof(None, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.sdfr_decoder_eval_lattice(stock.native().handle, lat.data_ptr(), 1, 40, sdf.data_ptr(), din.data_ptr(), _lib.MLP_TCGEN05, _lib.stream_ptr()))
    b.record(); torch.cuda.synchronize()
    prof(buf, 0)
    names = ["producer wait empty", "mma wait act(epilogue)", "mma wait full(weights)", "epi wait acc(mma)", "epi busy fwd", "epi busy bwd", "epi busy bwd-first"]
    print(f"run {it}: {a.elapsed_time(b):.3f} ms; CTA0 cycles:", {n: int(buf[i]) for i, n in enumerate(names)}, flush=True)
