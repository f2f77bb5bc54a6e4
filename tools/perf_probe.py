"""Times the two decoder launches of a refine iteration in isolation (coarse lattice pass, accurate
forward+gradient on a band-sized row list) and checks them against the FFMA kernel.
SDFR_LIB=<other build> python tools/perf_probe.py compares kernel variants (one process per build)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdflabel_b200 import _lib  # noqa: E402
from sdflabel_b200.deepsdf.workspace import setup_dsdf  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")
torch.manual_seed(0)
stock, L = setup_dsdf(os.path.join(ROOT, "assets", "deepsdf_synth.pt"), precision=torch.float32)
stock = stock.to(dev)
h = stock.native().handle
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SDFR_"))


def timeit(fn, reps=12):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[3:]))


lat = torch.nn.functional.normalize(torch.tensor([[0.5, 0.7, 0.5]]), dim=1).to(dev)
ng = 64000
sdf_ref = torch.empty(ng, device=dev)
sdf = torch.empty(ng, device=dev)
_lib.check(lib.sdfr_decoder_eval_lattice(h, lat.data_ptr(), 1, 40, sdf_ref.data_ptr(), 0, _lib.MLP_FFMA, _lib.stream_ptr()))


def coarse():
    _lib.check(lib.sdfr_decoder_eval_lattice(h, lat.data_ptr(), 1, 40, sdf.data_ptr(), 0, _lib.MLP_TCGEN05_COARSE,
                                             _lib.stream_ptr()))


t = timeit(coarse) if os.environ.get("PROBE_COARSE", "1") == "1" else 0.0
err = float((sdf - sdf_ref).abs().max())
extra = ""
print(f"[{tag}] coarse lattice 40^3: {t * 1e3:.1f} us, max |sdf - ffma| {err:.2e}{extra}", flush=True)

ns = [int(v) for v in os.environ.get("PROBE_N", "1850,7400").split(",")]
for n in ns:
    x = torch.cat([lat.cpu().expand(n, -1), torch.rand(n, 3) * 2 - 1], 1).contiguous().to(dev)
    outs = {}
    for name, impl in (("ffma", _lib.MLP_FFMA), ("tc", _lib.MLP_TCGEN05)):
        s_ = torch.empty(n, device=dev)
        d_ = torch.empty(n, L + 3, device=dev)

        def run():
            _lib.check(lib.sdfr_decoder_eval(h, x.data_ptr(), n, s_.data_ptr(), d_.data_ptr(), impl, _lib.stream_ptr()))
        tt = timeit(run)
        outs[name] = (s_.clone(), d_.clone(), tt)
    es = float((outs["tc"][0] - outs["ffma"][0]).abs().max())
    eg = float((outs["tc"][1] - outs["ffma"][1]).abs().max())
    print(f"[{tag}] accurate fwd+grad n={n}: tc {outs['tc'][2] * 1e3:.1f} us (ffma {outs['ffma'][2] * 1e3:.1f} us), "
          f"max diff sdf {es:.2e} grad {eg:.2e}", flush=True)
