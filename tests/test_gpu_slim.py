"""Channel slimming on the device (csnet_slim_gather): same result as the reference goldens with the model's tensors on the GPU, and the
slimmed model runs on the engine (eval forward + one training step)."""
import pytest
import torch

from sod100k_b200 import synth
from sod100k_b200.model import csnet
from sod100k_b200.trainer import Trainer
from tests import fixtures
from tests.test_slim import CASES, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag,thres", CASES)
def test_device_slimming_matches_the_reference(tag, thres):
    cfg, sd = fixtures.checkpoint(tag)
    m = csnet.CSNet(cfg)
    m.load_state_dict(sd)
    m.cuda()
    slimmed, new_cfg = check_against_golden(m, tag, thres, cfg)
    assert all(p.is_cuda for p in slimmed.parameters())
    x = torch.from_numpy(synth.randn_images(2, 64, 96, 3)).cuda()
    slimmed.eval()
    with torch.no_grad():
        y = slimmed(x)
    assert tuple(y.shape) == (2, 1, 64, 96) and torch.isfinite(y).all()
    slimmed.train()
    loss = Trainer(slimmed, lr=1e-4, weight_decay=5e-3).step(x, torch.from_numpy(synth.random_masks(2, 64, 96, 4)).cuda())
    assert torch.isfinite(loss)
