"""tests/golden/slim.npz: the reference's own channel slimming — `finetune_model` + `build_model_with_weight`
(/root/reference/CSNet_training/model/csnet.py:763-879), imported unmodified — on the two shipped checkpoints at thresholds that prune
a little and a lot.  Per case: the new layer_config and, for every tensor of the slimmed model's state_dict (in order), its shape, its
float64 sum of squares and 16 sampled elements (index stream default_rng(1000 + position)).  Harness-side shims only (the reference is untouched): `collections.Iterable` (removed in Python 3.10)
and a ragged-tolerant `np.array` (csnet.py:798 builds an array of three masks of different lengths, which numpy >= 1.24 refuses
without dtype=object).  Run in the build container:  python tests/golden/make_slim_golden.py"""
import collections, collections.abc, contextlib, io, json, os, sys

import numpy as np

collections.Iterable = collections.abc.Iterable
_np_array = np.array


def _ragged_ok(x, *a, **k):
    try:
        return _np_array(x, *a, **k)
    except ValueError:
        return _np_array(x, dtype=object)


np.array = _ragged_ok
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/CSNet_training")
import torch  # noqa: E402

import model.csnet as T  # noqa: E402  (the reference, training variant)
from tests import fixtures  # noqa: E402

CASES = [("csnet-L-x2", 1e-3), ("csnet-L-x2", 1e-2), ("csnet-L-x1", 3e-3)]
N_SAMPLE = 16


def cfg_json(cfg):
    return json.dumps([[np.asarray(e, np.float64).reshape(-1).tolist() for e in entry] for entry in cfg[:-1]] + [[int(s) for s in cfg[-1]]])


out = {}
for tag, thres in CASES:
    cfg, sd = fixtures.checkpoint(tag)
    with contextlib.redirect_stdout(io.StringIO()):
        m = T.CSNet(layer_config=cfg)
        m.load_state_dict(sd)
        new_cfg, masks = T.finetune_model(m, "/tmp", cfg, thres)
        slim = T.build_model_with_weight(new_cfg, m, masks)
    key = f"{tag}@{thres:g}"
    out[f"{key}/config"] = np.array(cfg_json(new_cfg))
    meta, ss, samples = {}, [], []
    for i, (k, v) in enumerate(slim.state_dict().items()):
        a = v.detach().cpu().numpy().astype(np.float64).reshape(-1)
        idx = np.random.default_rng(1000 + i).integers(0, max(a.size, 1), N_SAMPLE)
        ss.append(float((a * a).sum()))
        samples.append(a[idx] if a.size else np.zeros(N_SAMPLE))
        meta[k] = list(v.shape)
    out[f"{key}/shapes"] = np.array(json.dumps(meta))          # insertion order = state_dict order
    out[f"{key}/ss"] = np.array(ss)
    out[f"{key}/samples"] = np.stack(samples)
    print(key, "params", sum(p.numel() for p in slim.parameters()), "of", sum(p.numel() for p in m.parameters()))
np.savez_compressed(os.path.join(HERE, "slim.npz"), **out)
print("wrote", os.path.join(HERE, "slim.npz"), os.path.getsize(os.path.join(HERE, "slim.npz")), "bytes")
