"""Loaders for the committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py)."""
from __future__ import annotations

import functools
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cfg_from_json(text: str):
    """Inverse of make_golden.cfg_to_json: list of [in_split, out_split(, dil_split)] float64 arrays + stages."""
    d = json.loads(text)
    cfg = [[np.asarray(e, np.float64) for e in entry] for entry in d["blocks"]]
    cfg.append(list(d["stages"]))
    return cfg


@functools.lru_cache(maxsize=None)
def checkpoint(tag: str):
    """(layer_config, {key: torch tensor}) of a shipped checkpoint re-serialised as npz."""
    z = np.load(os.path.join(GOLDEN, f"{tag}.npz"))
    cfg = cfg_from_json(str(z["__layer_config__"]))
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k != "__layer_config__"}
    return cfg, sd


@functools.lru_cache(maxsize=None)
def forward_golden():
    z = np.load(os.path.join(GOLDEN, "forward.npz"))
    meta = json.loads(str(z["__meta__"]))
    return z, meta


def synthetic_model(tag: str):
    """(layer_config, state dict) of an un-pruned architecture with seeded synthetic weights."""
    from sod100k_b200 import synth

    _, meta = forward_golden()
    m = meta[tag]
    cfg = cfg_from_json(m["layer_config"])
    sd = synth.synth_state({k: tuple(v) for k, v in m["shapes"].items()}, m["seed"])
    return cfg, {k: torch.from_numpy(v) for k, v in sd.items()}, m
