"""Shared helpers of the parity tests (the oracle is the checker, never the thing under test)."""
import json
import os

import numpy as np
import torch

from oracle import prior as P
from oracle import sdf_oracle as O


def our_decoder_from_state(spec: O.DecoderSpec, state_dict, device="cuda"):
    """Instantiates the product Decoder exactly the way setup_dsdf does."""
    from sdflabel_b200.deepsdf.networks.deep_sdf_decoder_scale import Decoder
    ns = spec.to_json()["NetworkSpecs"]
    dec = Decoder(spec.latent_size, **ns)
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
    dec.load_state_dict(sd)
    return dec.to(device).eval()


def golden_decoder(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"decoder_{name}.npz"))
    spec = O.DecoderSpec.from_json(json.loads(bytes(g["spec_json"]).decode()))
    sd = {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd::")}
    return g, spec, sd


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(1e-30, np.abs(b).max())


def frac_within(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ok = np.abs(a - b) <= atol + rtol * np.abs(b)
    return float(ok.mean())
