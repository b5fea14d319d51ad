"""Device SalMetric (sod100k_b200/salmetric.py, csnet_salmetric_hist) against the numpy restatement in oracle/salmetric.py,
which tests/test_salmetric.py pins to the reference's C++ loops."""
import numpy as np
import pytest
import torch

from oracle import salmetric as O
from sod100k_b200 import salmetric, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,seed", [((5, 1, 64, 96), 3), ((2, 224, 224), 4), ((1, 1, 7, 5), 5)])
def test_device_histograms_reproduce_the_reference_metric(shape, seed):
    rng = np.random.default_rng(seed)
    prob = rng.random(shape, dtype=np.float32)
    prob.reshape(-1)[:4] = [0.0, 1.0, 128 / 255, 0.99999994]                # bin edges
    gt = (rng.random(shape) > 0.7).astype(np.uint8) * 255
    gt.reshape(shape[0], -1)[-1, :] = 0                                     # an image without foreground
    gt.reshape(-1)[5:8] = [128, 129, 127]                                   # around the > 128 rule
    m = salmetric.SalMetric()
    half = max(1, shape[0] // 2)                                            # two updates: the evaluator streams
    m.update(torch.from_numpy(prob[:half]).cuda(), torch.from_numpy(gt[:half]).cuda())
    if half < shape[0]:
        m.update(torch.from_numpy(prob[half:]).cuda(), torch.from_numpy(gt[half:]).cuda())
    got = m.compute()
    sal = [O.quantise(p.reshape(-1, p.shape[-1])) for p in prob]
    ref = O.evaluate(sal, [g.reshape(-1, g.shape[-1]) for g in gt])
    assert got["argmax"] == ref["argmax"]
    for key in ("max_f", "mean_f", "precision", "recall", "mean_precision", "mean_recall", "mae"):
        assert got[key] == pytest.approx(ref[key], rel=2e-6, abs=2e-7), key


def test_on_network_output_of_the_blob_set():
    """End to end: engine logits -> sigmoid -> device metric == oracle metric of the same maps."""
    from sod100k_b200 import checkpoints
    model, cfg, _ = checkpoints.build_from_npz("csnet-L-x2")
    model.cuda().eval()
    xb, masks = synth.blob_images(4, 224, 224, 1235)
    with torch.no_grad():
        prob = torch.sigmoid(model(torch.from_numpy(xb).cuda()))
    gt = torch.from_numpy((masks * 255).astype(np.uint8)).cuda()
    m = salmetric.SalMetric()
    m.update(prob, gt)
    got = m.compute()
    ref = O.evaluate([O.quantise(p[0]) for p in prob.cpu().numpy()], [g[0] for g in (masks * 255).astype(np.uint8)])
    assert got["max_f"] == pytest.approx(ref["max_f"], abs=1e-6) and got["mae"] == pytest.approx(ref["mae"], abs=1e-6)


def test_rejects_cpu_tensors():
    with pytest.raises(Exception):
        salmetric.SalMetric().update(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4, dtype=torch.uint8))
