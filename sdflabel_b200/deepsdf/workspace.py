"""DeepSDF checkpoint loader.

Mirror of ``setup_dsdf`` / ``convert_to_precision`` in the reference's
sdfrenderer/deepsdf/workspace.py:167-195: ``<path>.json`` holds the network spec,
``<path>.pt`` holds ``{"epoch", "model_state_dict"}`` saved from a DataParallel
wrapper (keys prefixed with ``module.``).  Only the loader is on the hot path; the
rest of the reference's workspace helpers (DeepSDF experiment directories) is not.
"""
from __future__ import annotations

import json
import os

import torch
import torch.nn as nn


def setup_dsdf(dir, mode='eval', precision=torch.float16):
    specs_filename = os.path.splitext(dir)[0] + '.json'
    if not os.path.isfile(specs_filename):
        raise Exception('The experiment directory does not include specifications file "specs.json"')
    with open(specs_filename) as f:
        specs = json.load(f)
    arch = __import__("sdflabel_b200.deepsdf.networks." + specs["NetworkArch"], fromlist=["Decoder"])
    latent_size = specs["CodeLength"]
    net_specs = dict(specs["NetworkSpecs"])
    net_specs.pop('samples_per_scene', None)  # single-model scale head, as in the reference
    decoder = arch.Decoder(latent_size, **net_specs)
    try:
        saved = torch.load(dir, map_location='cpu')
    except Exception:
        saved = torch.load(dir, map_location='cpu', weights_only=False)
    state = saved["model_state_dict"]
    state = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state.items()}
    decoder.load_state_dict(state)
    decoder.saved_epoch = saved.get("epoch")
    convert_to_precision(decoder, precision)
    if mode == 'train':
        decoder.train()
    elif mode == 'eval':
        decoder.eval()
    return decoder, latent_size


def convert_to_precision(model, precision):
    """The kernels always compute in fp32 (>= the reference's fp16 default); the
    cast only sets the dtype of the tensors the module exchanges with its caller."""
    model.to(dtype=precision)
    for layer in model.modules():
        if isinstance(layer, nn.BatchNorm2d):
            layer.float()
