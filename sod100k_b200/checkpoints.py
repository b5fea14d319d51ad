"""Load the .npz re-serialisation of a CSNet checkpoint (tests/golden/*.npz: the shipped
csnet-L-x{1,2} weights + layer_config, produced by tests/golden/make_golden.py) without the reference."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def layer_config_from_json(text: str):
    d = json.loads(text)
    cfg = [[np.asarray(e, np.float64) for e in entry] for entry in d["blocks"]]
    cfg.append(list(d["stages"]))
    return cfg


def load_npz(tag_or_path: str):
    """Returns (layer_config, {state_dict key: numpy array})."""
    path = tag_or_path if os.path.exists(tag_or_path) else os.path.join(GOLDEN_DIR, f"{tag_or_path}.npz")
    z = np.load(path)
    cfg = layer_config_from_json(str(z["__layer_config__"]))
    return cfg, {k: z[k] for k in z.files if k != "__layer_config__"}


def build_from_npz(tag_or_path: str):
    """CSNet module with the checkpoint loaded (CPU; call .cuda())."""
    import torch

    from .model import csnet

    cfg, sd = load_npz(tag_or_path)
    m = csnet.CSNet(cfg)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return m, cfg, sd
