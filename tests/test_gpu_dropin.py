"""The call sequence of the reference's test.py (CSNet/test.py:35-99) against the drop-in `model` package: the scripts
themselves need yacs / skimage / a dataset, so the test replays their model-facing calls in order."""
import contextlib
import importlib
import io
import os
import sys

import numpy as np
import pytest
import torch

from oracle import csnet_oracle as O
from sod100k_b200 import synth
from tests import fixtures

pytestmark = pytest.mark.gpu


def test_test_py_call_sequence(tmp_path):
    pkg_parent = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sod100k_b200")
    sys.path.insert(0, pkg_parent)                        # `model` now resolves to sod100k_b200/model
    try:
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        model_lib = importlib.import_module("model." + "csnet")                    # test.py:37
        from model.utils.simplesum_octconv import simplesum                         # test.py:12
        cfg, sd = fixtures.checkpoint("csnet-L-x2")
        model_lib.save_layer_config(cfg, str(tmp_path), 0)
        model = model_lib.build_model(predefine=str(tmp_path / "layer_config_0.bin"))   # test.py:39
        model.cuda()                                                                  # test.py:40
        with contextlib.redirect_stdout(io.StringIO()):
            prams, flops = simplesum(model, inputsize=(3, 224, 224), device=0)        # test.py:41
        assert abs(prams / 1e6 - 0.1409) < 1e-4 and abs(flops / 1e9 - 0.7167) < 1e-4
        torch.save({"epoch": 20, "arch": "csnet", "state_dict": sd, "optimizer": {}}, tmp_path / "ck.pth.tar")
        checkpoint = torch.load(tmp_path / "ck.pth.tar", weights_only=False)           # test.py:47
        model.load_state_dict(checkpoint["state_dict"])                               # test.py:49
        model.eval()                                                                  # test.py:59
        img = synth.randn_images(1, 224, 224, 4)[0]                                   # stands in for the normalised image
        with torch.no_grad():                                                         # test.py:70
            input_var = torch.unsqueeze(torch.FloatTensor(img), 0).cuda()             # test.py:86-89
            predict = model(input_var)                                                # test.py:90
            predict = torch.sigmoid(predict[0].squeeze(0).squeeze(0)).data.cpu().numpy()   # test.py:91-93
            ref = torch.sigmoid(O.csnet_forward(cfg, sd, torch.from_numpy(img[None])))[0, 0].numpy()
        assert np.abs(predict - ref).max() <= 1e-3
        png = (predict * 255).astype(np.uint8)                                        # test.py:94-96 (without the resize)
        assert np.abs(png.astype(int) - (ref * 255).astype(np.uint8).astype(int)).max() <= 1
    finally:
        sys.path.remove(pkg_parent)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
