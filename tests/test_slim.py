"""Channel slimming against the reference's own finetune_model + build_model_with_weight, run unmodified in the build
container by tests/golden/make_slim_golden.py (tests/golden/slim.npz): the new layer_config and every tensor of the slimmed state_dict."""
import json
import os

import numpy as np
import pytest
import torch

from sod100k_b200 import slim
from sod100k_b200.model import csnet
from tests import fixtures

CASES = [("csnet-L-x2", 1e-3), ("csnet-L-x2", 1e-2), ("csnet-L-x1", 3e-3)]


def _golden():
    return np.load(os.path.join(fixtures.GOLDEN, "slim.npz"))


def _cfg_lists(cfg):
    return [[np.asarray(e, np.float64).reshape(-1).tolist() for e in entry] for entry in cfg[:-1]] + [[int(s) for s in cfg[-1]]]


def check_against_golden(model, tag, thres, base_cfg):
    z = _golden()
    key = f"{tag}@{thres:g}"
    new_cfg, masks = slim.finetune_config(model, base_cfg, thres)
    assert _cfg_lists(new_cfg) == json.loads(str(z[f"{key}/config"]))
    slimmed = slim.build_model_with_weight(new_cfg, model, masks)
    shapes = json.loads(str(z[f"{key}/shapes"]))
    sd = slimmed.state_dict()
    assert list(sd.keys()) == list(shapes.keys())
    for i, (k, v) in enumerate(sd.items()):
        assert list(v.shape) == shapes[k], k
        a = v.detach().cpu().numpy().astype(np.float64).reshape(-1)
        idx = np.random.default_rng(1000 + i).integers(0, max(a.size, 1), 16)
        assert abs((a * a).sum() - z[f"{key}/ss"][i]) <= 1e-9 * max(1.0, z[f"{key}/ss"][i]), k
        if a.size:
            assert np.array_equal(a[idx], z[f"{key}/samples"][i]), k          # copies: bit-exact
    return slimmed, new_cfg


@pytest.mark.parametrize("tag,thres", CASES)
def test_slimming_matches_the_reference(tag, thres):
    cfg, sd = fixtures.checkpoint(tag)
    m = csnet.CSNet(cfg)
    m.load_state_dict(sd)
    slimmed, new_cfg = check_against_golden(m, tag, thres, cfg)
    assert sum(p.numel() for p in slimmed.parameters()) < sum(p.numel() for p in m.parameters())


def test_build_model_finetune_arguments(tmp_path):
    """build_model(finetune=True, ...) as CSNet_training/finetune.py calls it: config pickle in, slimmed config pickle out
    (layer_config_finetune_<epoch>.bin), weights carried over with load_weight='FINETUNE'."""
    cfg, sd = fixtures.checkpoint("csnet-L-x2")
    m = csnet.CSNet(cfg)
    m.load_state_dict(sd)
    csnet.save_layer_config(cfg, str(tmp_path), 0, latest=True)
    pre = os.path.join(str(tmp_path), "layer_config_latest.bin")
    new = csnet.build_model(epoch=3, predefine=pre, save_path=str(tmp_path), model=m, load_weight="FINETUNE", finetune_thres=1e-2, finetune=True)
    saved = csnet.load_layer_config(os.path.join(str(tmp_path), "layer_config_finetune_3.bin"))
    z = _golden()
    assert _cfg_lists(saved) == json.loads(str(z["csnet-L-x2@0.01/config"]))
    assert sum(p.numel() for p in new.parameters()) == 55740
    fresh = csnet.build_model(epoch=3, predefine=pre, save_path=str(tmp_path), model=m, load_weight="NO", finetune_thres=1e-2, finetune=True)
    assert [tuple(p.shape) for p in fresh.parameters()] == [tuple(p.shape) for p in new.parameters()]
    with pytest.raises(NotImplementedError):
        csnet.build_model(epoch=2, predefine="does-not-exist")
