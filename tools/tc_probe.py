"""GPU probe of the tcgen05 MLP kernel against the FFMA kernel (run under a short timeout)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sdflabel_b200 import _lib
from sdflabel_b200.deepsdf.networks.deep_sdf_decoder_scale import Decoder
from sdflabel_b200.deepsdf.workspace import setup_dsdf

lib = _lib.load()
dev = torch.device("cuda")
torch.manual_seed(0)


def compare(dec, n, tag, grad=True):
    L = dec.latent_size
    x = torch.cat([torch.nn.functional.normalize(torch.randn(n, L), dim=1), torch.rand(n, 3) * 2 - 1], 1).to(dev)
    outs = {}
    for name, impl in (("ffma", _lib.MLP_FFMA), ("tc", _lib.MLP_TCGEN05)):
        sdf = torch.full((n,), float("nan"), device=dev)
        din = torch.full((n, L + 3), float("nan"), device=dev) if grad else None
        _lib.check(lib.sdfr_decoder_eval(dec.native().handle, x.data_ptr(), n, sdf.data_ptr(), _lib.ptr(din), impl,
                                         _lib.stream_ptr()))
        torch.cuda.synchronize()
        outs[name] = (sdf.cpu().numpy(), din.cpu().numpy() if grad else None)
    ds = np.abs(outs["ffma"][0] - outs["tc"][0])
    msg = f"{tag:28s} n={n:6d} sdf max diff {np.nanmax(ds):.3e} nan {int(np.isnan(outs['tc'][0]).sum())}"
    if grad:
        dg = np.abs(outs["ffma"][1] - outs["tc"][1])
        msg += f" | grad max diff {np.nanmax(dg):.3e} (|g| max {np.abs(outs['ffma'][1]).max():.3e}) nan {int(np.isnan(outs['tc'][1]).sum())}"
    print(msg, flush=True)
    return outs


print("caps", lib.sdfr_caps(), flush=True)
tiny = Decoder(3, [64]).to(dev).eval()
print("tiny tc ok:", tiny.native().tcgen05, flush=True)
compare(tiny, 64, "tiny 6-64-1 fwd only", grad=False)
compare(tiny, 64, "tiny 6-64-1")
compare(tiny, 1000, "tiny 6-64-1")
mid = Decoder(3, [128, 128, 128], latent_in=(2,), norm_layers=(0, 1, 2), weight_norm=True).to(dev).eval()
compare(mid, 500, "mid 3x128 skip")
wide = Decoder(3, [512, 512], norm_layers=(0, 1), weight_norm=True).to(dev).eval()
compare(wide, 500, "wide 2x512")
stock, L = setup_dsdf(os.path.join(ROOT, "assets", "deepsdf_synth.pt"), precision=torch.float32)
stock = stock.to(dev)
print("stock tc ok:", stock.native().tcgen05, flush=True)
compare(stock, 64, "stock fwd only", grad=False)
compare(stock, 64, "stock")
compare(stock, 64000, "stock")
# timing on the lattice
lat = torch.nn.functional.normalize(torch.tensor([[0.5, 0.7, 0.5]]), dim=1).to(dev)
sdf = torch.empty(64000, device=dev)
din = torch.empty(64000, 6, device=dev)
for name, impl in (("ffma", _lib.MLP_FFMA), ("tc", _lib.MLP_TCGEN05)):
    for withgrad in (False, True):
        ts = []
        for i in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.sdfr_decoder_eval_lattice(stock.native().handle, lat.data_ptr(), 1, 40, sdf.data_ptr(),
                                                     din.data_ptr() if withgrad else 0, impl, _lib.stream_ptr()))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"lattice 40^3 {name:5s} grad={withgrad}: {np.median(ts[2:]):.3f} ms", flush=True)
print("probe done", flush=True)
