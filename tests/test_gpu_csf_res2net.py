"""CSF+Res2Net (config 5) on the GPU: backbone on torch/cuDNN (library), CSF head on the engine."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import csf_res2net_oracle as R
from sod100k_b200 import compiler_r, runtime, synth
from sod100k_b200.networks import csf_res2net
from tests import fixtures

pytestmark = pytest.mark.gpu


def _golden():
    z = np.load(os.path.join(fixtures.GOLDEN, "csf_res2net.npz"))
    return z, json.loads(str(z["__meta__"]))


def _model(meta):
    m = csf_res2net.build_model()
    sd = synth.synth_state_r({k: tuple(v) for k, v in meta["shapes"].items()}, meta["seed"])
    res = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert list(m.state_dict().keys()) == list(meta["shapes"].keys())
    return m.cuda().eval(), {k: torch.from_numpy(v) for k, v in sd.items()}


def test_fp32_matches_reference_goldens_and_head_taps():
    z, meta = _golden()
    m, sd = _model(meta)
    for tag, (h, w, seed) in meta["cases"].items():
        x = torch.from_numpy(synth.randn_images(1, h, w, seed))
        with torch.no_grad():
            y = m(x.cuda()).cpu().numpy()
        ref = z[f"{tag}/logits"]
        assert np.abs(y - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()), tag
    # head in isolation on the oracle's (CPU, fp32) backbone features: generic kernels, then the tensor-core MIX kernel
    h, w, seed = meta["cases"]["a"]
    x = torch.from_numpy(synth.randn_images(2, h, w, seed))
    taps = {}
    with torch.no_grad():
        ref = R.csfnet_forward(sd, x, taps)
    feats = [f.cuda().contiguous() for f in taps["feats"]]
    for tc in (False, True):
        prog = compiler_r.compile_csf_head(sd, [tuple(f.shape[1:]) for f in feats], h, w, "fp32", reuse_arena=False, tensor_core=tc)
        plan = runtime.Plan(prog, max_batch=2)
        y = torch.empty((2, 1, h, w), device="cuda")
        plan.run(2, [f.data_ptr() for f in feats] + [y.data_ptr()], torch.cuda.current_stream().cuda_stream)
        assert (y.cpu() - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
        for name, r in (("fuse/1", taps["fuse"][1]), ("ms/3", taps["ms"][3]), ("fuse1x1/0", taps["fuse1x1"])):
            got = plan.read_tensor(prog.taps[name], 2).cpu()
            assert (got - r).abs().max().item() <= 1e-3 * max(1.0, r.abs().max().item()), (tc, name)


def test_fp16_head_on_tensor_cores():
    z, meta = _golden()
    m, sd = _model(meta)
    m.set_precision("fp16")
    h, w, seed = meta["cases"]["b"]
    x = torch.from_numpy(synth.randn_images(2, h, w, seed))
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        ref = R.csfnet_forward(sd, x)
    # fp16 backbone (cuDNN autocast) + fp16 head storage: stated tolerance 3e-2 of the logit range
    assert (y - ref).abs().max().item() <= 3e-2 * max(1.0, ref.abs().max().item())
    assert (torch.sigmoid(y) - torch.sigmoid(ref)).abs().max().item() <= 2e-2
