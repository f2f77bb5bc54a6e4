"""Band pass (accurate forward + input gradient on a short row list) on CTA pairs against the single-CTA
16-point kernel: `python tools/band_probe.py` runs itself twice (SDFR_BAND_PAIR=0 / 1), times both and compares
sdf / dinput bit for bit.  Row counts: PROBE_N (default 1850,1813,37,3700)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NS = [int(v) for v in os.environ.get("PROBE_N", "1850,1813,37,3700").split(",")]


def child(tag):
    import torch
    from sdflabel_b200 import _lib
    from sdflabel_b200.deepsdf.workspace import setup_dsdf
    lib = _lib.load()
    dev = torch.device("cuda")
    dec, L = setup_dsdf(os.path.join(ROOT, "assets", "deepsdf_synth.pt"), precision=torch.float32)
    dec = dec.to(dev)
    h = dec.native().handle
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for n in NS:
        g = torch.Generator().manual_seed(n)
        lat = torch.nn.functional.normalize(torch.tensor([[0.5, 0.7, 0.5]]), dim=1)
        x = torch.cat([lat.expand(n, -1), torch.rand(n, 3, generator=g) * 2 - 1], 1).contiguous().to(dev)
        s_ = torch.empty(n, device=dev)
        d_ = torch.empty(n, L + 3, device=dev)
        f_ = torch.empty(n, device=dev)

        def run(grad=True):
            _lib.check(lib.sdfr_decoder_eval(h, x.data_ptr(), n, (s_ if grad else f_).data_ptr(), d_.data_ptr() if grad else 0,
                                             _lib.MLP_TCGEN05, _lib.stream_ptr()))
        ts = []
        for i in range(12):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        run(False)
        torch.cuda.synchronize()
        sr = torch.empty(n, device=dev); dr = torch.empty(n, L + 3, device=dev)
        _lib.check(lib.sdfr_decoder_eval(h, x.data_ptr(), n, sr.data_ptr(), dr.data_ptr(), _lib.MLP_FFMA, _lib.stream_ptr()))
        torch.cuda.synchronize()
        print(f"[{tag}] n={n}: {np.median(ts[3:]) * 1e3:.1f} us; vs ffma: sdf {float((s_ - sr).abs().max()):.2e} "
              f"grad {float((d_ - dr).abs().max()):.2e}; fwd-only == fwd+grad sdf: {bool(torch.equal(s_, f_))}", flush=True)
        out[f"sdf_{n}"] = s_.cpu().numpy(); out[f"din_{n}"] = d_.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", f"band_probe_{tag}.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    for tag, v in (("single", "0"), ("pair", "1")):
        env = dict(os.environ, SDFR_BAND_PAIR=v)
        r = subprocess.run([sys.executable, __file__, tag], env=env, timeout=600)
        if r.returncode:
            print(f"{tag}: rc={r.returncode}")
            sys.exit(1)
    a = np.load(os.path.join(ROOT, "gpurun_out", "band_probe_single.npz"))
    b = np.load(os.path.join(ROOT, "gpurun_out", "band_probe_pair.npz"))
    ok = True
    for k in a.files:
        same = np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
        ok &= same
        print(f"{k}: bit-identical {same}" + ("" if same else f" (max diff {np.abs(a[k] - b[k]).max():.3e})"))
    print("BIT-IDENTICAL" if ok else "DIFFERENT")
