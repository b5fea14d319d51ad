"""CSF+Res2Net (config 5), CPU side: oracle vs reference goldens; head compiler + kernel bodies (host emulation) vs oracle."""
import json
import os

import numpy as np
import torch

from oracle import csf_res2net_oracle as R
from sod100k_b200 import compiler_r, synth
from tests import emu, fixtures


def _golden():
    z = np.load(os.path.join(fixtures.GOLDEN, "csf_res2net.npz"))
    return z, json.loads(str(z["__meta__"]))


def _state(meta):
    sd = synth.synth_state_r({k: tuple(v) for k, v in meta["shapes"].items()}, meta["seed"])
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def test_oracle_matches_reference_goldens():
    z, meta = _golden()
    sd = _state(meta)
    assert sum(v.numel() for k, v in sd.items() if "running" not in k and "num_batches" not in k) == meta["params"]
    for tag, (h, w, seed) in meta["cases"].items():
        with torch.no_grad():
            y = R.csfnet_forward(sd, torch.from_numpy(synth.randn_images(1, h, w, seed))).numpy()
        assert np.abs(y - z[f"{tag}/logits"]).max() <= 2e-5 * max(1.0, np.abs(y).max())


def test_emulated_head_program_matches_oracle():
    z, meta = _golden()
    sd = _state(meta)
    h, w, seed = meta["cases"]["a"]
    x = torch.from_numpy(synth.randn_images(1, h, w, seed))
    taps = {}
    with torch.no_grad():
        ref = R.csfnet_forward(sd, x, taps).numpy()
    feats = [np.ascontiguousarray(f.numpy()) for f in taps["feats"]]
    prog = compiler_r.compile_csf_head(sd, [f.shape[1:] for f in feats], h, w, "fp32", reuse_arena=False)
    y = np.zeros((1, 1, h, w), np.float32)
    got = emu.run_ext(prog, feats + [y], 1, taps=["fuse/0", "fuse/3", "ms/1", "fuse1x1/0"])
    for name, r in (("fuse/0", taps["fuse"][0]), ("fuse/3", taps["fuse"][3]), ("ms/1", taps["ms"][1]), ("fuse1x1/0", taps["fuse1x1"])):
        r = r.numpy()
        assert np.abs(got[name] - r).max() <= 2e-4 * max(1.0, np.abs(r).max()), name
    assert np.abs(y - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())
    assert np.abs(y - z["a/logits"]).max() <= 2e-4 * max(1.0, np.abs(ref).max())
