"""In-kernel phase timers of mlp_tc_kernel on short row lists (library built with SDFR_NVCC_FLAGS=-DSDFR_TC_PROFILE)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdflabel_b200 import _lib
if os.environ.get("SDFR_LIB"):
    _lib.LIB_PATH = os.environ["SDFR_LIB"]
from sdflabel_b200.deepsdf.workspace import setup_dsdf
lib = _lib.load()
dev = torch.device("cuda")
stock, L = setup_dsdf(os.path.join(ROOT, "assets", "deepsdf_synth.pt"), precision=torch.float32)
stock = stock.to(dev)
prof = lib.sdfr_debug_tc_prof
prof.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
buf = (ctypes.c_ulonglong * 16)()
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SDFR_TC"))
for n in [int(v) for v in os.environ.get("PROBE_N", "64,1850").split(",")]:
    x = torch.rand(n, L + 3, device=dev) * 2 - 1
    s_ = torch.empty(n, device=dev); d_ = torch.empty(n, L + 3, device=dev)
    for it in range(3):
        prof(None, 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.sdfr_decoder_eval(stock.native().handle, x.data_ptr(), n, s_.data_ptr(), d_.data_ptr(), _lib.MLP_TCGEN05, _lib.stream_ptr()))
        b.record(); torch.cuda.synchronize()
        prof(buf, 0)
    names = ["producer wait empty", "mma wait act(epilogue)", "mma wait full(weights)", "epi wait acc(mma)", "epi busy fwd", "epi busy bwd", "epi busy bwd-first"]
    print(f"[{tag}] n={n}: {a.elapsed_time(b) * 1e3:.1f} us; CTA0 cycles:", {nm: int(buf[i]) for i, nm in enumerate(names)}, flush=True)
