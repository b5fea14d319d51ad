"""Generate the golden fixtures by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Needs /root/reference (absent on the GPU box — the fixtures are what travels).  The reference module
is imported as-is with one harness-side shim (`collections.Iterable`, removed in Python 3.10, is used
by /root/reference/CSNet/model/conv2d.py:15).  Nothing is copied from the reference: the fixtures hold
(a) the shipped checkpoints' tensors re-serialised as .npz (the function to match is defined by them),
(b) outputs of the reference forward on seeded synthetic inputs (sod100k_b200/synth.py).
"""
from __future__ import annotations

import collections
import collections.abc
import contextlib
import io
import json
import os
import sys

import numpy as np

collections.Iterable = collections.abc.Iterable          # harness-side shim, reference untouched
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/CSNet")

import torch  # noqa: E402

import model.csnet as ref  # noqa: E402  (the reference)
from sod100k_b200 import synth  # noqa: E402

CK = "/root/reference/CSNet/checkpoints"
N_SAMPLE = 64


def cfg_to_json(layer_config):
    out = []
    for entry in layer_config[:-1]:
        out.append([np.asarray(e, np.float64).tolist() if not np.isscalar(e) else [float(e)] for e in entry])
    return json.dumps(dict(blocks=out, stages=[int(s) for s in layer_config[-1]]))


def build(layer_config=None, predefine=None, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        if predefine:
            return ref.build_model(predefine=predefine)
        return ref.build_model(**kw)


def sample_idx(numel, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, numel, N_SAMPLE)


def run_case(model, x, taps_out=None):
    """Reference forward in eval mode; optionally record per-ILBlock / head taps via forward hooks."""
    hooks = []
    if taps_out is not None:
        for name, m in model.named_modules():
            if isinstance(m, ref.ILBlock) or name in ("oct_fuse.fuse", "oct_fuse.ms", "oct_fuse.fuse1x1"):
                def fn(mod, inp, out, name=name):
                    for b, t in enumerate(out if isinstance(out, (list, tuple)) else [out]):
                        if t is None:
                            continue
                        flat = t.detach().reshape(-1).numpy()
                        idx = sample_idx(flat.size, 99)
                        taps_out[f"{name}/{b}"] = np.concatenate(
                            [[flat.mean(), flat.std(), np.abs(flat).max()], flat[idx]]).astype(np.float32)
                hooks.append(m.register_forward_hook(fn))
    with torch.no_grad():
        y = model(torch.from_numpy(x))
    for h in hooks:
        h.remove()
    return y.numpy()


def main():
    torch.manual_seed(0)
    fwd = {}
    meta = {}
    for tag in ("csnet-L-x2", "csnet-L-x1"):
        model = build(predefine=f"{CK}/{tag}/{tag}.bin")
        ck = torch.load(f"{CK}/{tag}/{tag}.pth.tar", map_location="cpu", weights_only=False)
        model.load_state_dict(ck["state_dict"])
        model.eval()
        arrays = {k: v.numpy() for k, v in ck["state_dict"].items()}
        np.savez(os.path.join(HERE, f"{tag}.npz"), __layer_config__=np.array(cfg_to_json(model.layer_config)), **arrays)
        # 224x224, N=2, noise + blobs (full logits)
        x = synth.randn_images(2, 224, 224, 1234)
        taps = {}
        fwd[f"{tag}/randn224"] = run_case(model, x, taps if tag.endswith("x2") else None)
        for k, v in taps.items():
            fwd[f"{tag}/randn224/tap/{k}"] = v
        xb, _ = synth.blob_images(2, 224, 224, 1235)
        fwd[f"{tag}/blobs224"] = run_case(model, xb)
        # non-square, N=1
        fwd[f"{tag}/randn96x160"] = run_case(model, synth.randn_images(1, 96, 160, 1237))
        # 512x512, N=1 (sampled)
        y = run_case(model, synth.randn_images(1, 512, 512, 1238)).reshape(-1)
        idx = np.random.default_rng(5).integers(0, y.size, 8192)
        fwd[f"{tag}/randn512/sample"] = y[idx]
        fwd[f"{tag}/randn512/stats"] = np.array([y.mean(), y.std(), y.min(), y.max()], np.float32)
        meta[tag] = dict(params=int(sum(p.numel() for p in model.parameters())), n_state=len(arrays))

    # un-pruned architectures with seeded synthetic weights (weights are regenerated from the seed by the tests)
    for tag, kw, seed, hw in (("init-x2", dict(basic_split=[0.5, 0.5], expand=2.0), 7, (224, 224)),
                              ("init-std", dict(basic_split=[1]), 8, (64, 64)),
                              ("init-3br", dict(basic_split=[0.5, 0.25, 0.25]), 9, (128, 128))):
        model = build(**kw)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        sd = synth.synth_state(shapes, seed)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        model.eval()
        taps = {}
        fwd[f"{tag}/randn"] = run_case(model, synth.randn_images(1, hw[0], hw[1], 1240 + seed), taps)
        for k, v in taps.items():
            fwd[f"{tag}/randn/tap/{k}"] = v
        meta[tag] = dict(layer_config=cfg_to_json(model.layer_config), seed=seed, hw=hw,
                         shapes={k: list(v) for k, v in shapes.items()}, kw={k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
    # same seed -> same initial parameters (construction order + initialisers are part of the surface)
    torch.manual_seed(0)
    model = build(basic_split=[0.5, 0.5], expand=2.0)
    for k, v in model.state_dict().items():
        if k in ("stage0.0.conv1x1.conv.weight", "stage2.1.conv3x3_2.convs.1.weight", "oct_fuse.fuse.conv.weight",
                 "oct_fuse.ms.convs.1.msconv.3.weight", "cls_layer.weight", "cls_layer.bias"):
            fwd[f"seed0-init/{k}"] = v.numpy().reshape(-1)[:256].copy()
    meta["torch"] = torch.__version__
    meta["numpy"] = np.__version__
    meta["threads"] = torch.get_num_threads()
    np.savez_compressed(os.path.join(HERE, "forward.npz"), __meta__=np.array(json.dumps(meta)), **fwd)
    print("wrote", sorted(os.listdir(HERE)))


def main_train():
    """Train-mode pin: the UNMODIFIED reference module in train mode with the dynamic-weight-decay hooks
    (flops_hook(expandflop=1.0), csnet.py:332-355), mean BCE-with-logits + WEIGHT * get_flops() (train.py:209-213, WEIGHT = 3.0
    as in configs/csnet-L-x2_train.yml), autograd.  Stored per checkpoint: loss, regulariser, for EVERY parameter the gradient's
    L2 norm + 16 sampled elements, the updated BN running statistics (8 sampled elements each)."""
    import torch.nn.functional as F

    out = {}
    meta = {"flops_weight": 3.0, "expandflop": 1.0, "n": 2, "hw": [64, 96], "seed": 71, "n_grad_samples": 16, "n_stat_samples": 8}
    for tag in ("csnet-L-x2", "csnet-L-x1"):
        model = build(predefine=f"{CK}/{tag}/{tag}.bin")
        ck = torch.load(f"{CK}/{tag}/{tag}.pth.tar", map_location="cpu", weights_only=False)
        model.load_state_dict(ck["state_dict"])
        model.train()
        model.flops_hook(expandflop=1.0)
        x = torch.from_numpy(synth.randn_images(2, 64, 96, 71))
        t = torch.from_numpy(synth.random_masks(2, 64, 96, 72))
        model.clear_flops()
        model.set_batchsize(2)
        y = model(x)
        loss = F.binary_cross_entropy_with_logits(y, t)
        reg = model.get_flops()
        total = loss + 3.0 * reg
        model.zero_grad()
        total.backward()
        out[f"{tag}/loss"] = np.array([loss.item()], np.float64)
        out[f"{tag}/reg"] = np.array([float(reg)], np.float64)
        out[f"{tag}/logits_sample"] = y.detach().reshape(-1).numpy()[sample_idx(y.numel(), 7)]
        for k, p_ in model.named_parameters():
            g = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).reshape(-1).double().numpy()
            idx = np.random.default_rng(abs(hash(k)) % (2 ** 31) if False else sum(map(ord, k))).integers(0, g.size, 16)
            out[f"{tag}/grad/{k}"] = np.concatenate([[np.sqrt((g * g).sum())], g[idx]])
        for k, b in model.named_buffers():
            if k.endswith("running_mean") or k.endswith("running_var"):
                v = b.reshape(-1).double().numpy()
                idx = np.random.default_rng(sum(map(ord, k))).integers(0, v.size, 8)
                out[f"{tag}/stat/{k}"] = v[idx]
    meta["torch"] = torch.__version__
    np.savez_compressed(os.path.join(HERE, "train.npz"), __meta__=np.array(json.dumps(meta)), **out)
    print("wrote train.npz", {k: v.tolist() for k, v in out.items() if k.endswith("/loss") or k.endswith("/reg")})


def main_r():
    """CSF+Res2Net (config 5): reference `networks.csf_res2net.build_model()` with seeded synthetic weights."""
    for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference/CSF+Res2Net")
    from networks.csf_res2net import build_model as build_r

    with contextlib.redirect_stdout(io.StringIO()):
        net = build_r().eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    sd = synth.synth_state_r(shapes, 21)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    out, feats = {}, {}
    def grab(mod, inp, o):
        feats["f"] = [t.detach().numpy() for t in o]        # (returns None: a hook's return value would replace the output)

    hooks = [net.base.register_forward_hook(grab)]
    for tag, hw, seed in (("a", (64, 96), 1301), ("b", (96, 96), 1302)):
        feats.clear()
        with torch.no_grad():
            y = net(torch.from_numpy(synth.randn_images(1, hw[0], hw[1], seed))).numpy()
        out[f"{tag}/logits"] = y
        for i, f in enumerate(feats["f"]):
            out[f"{tag}/feat{i}/stats"] = np.array([f.mean(), f.std(), np.abs(f).max()], np.float32)
    for h in hooks:
        h.remove()
    meta = dict(seed=21, shapes={k: list(v) for k, v in shapes.items()}, params=int(sum(p.numel() for p in net.parameters())),
                cases=dict(a=[64, 96, 1301], b=[96, 96, 1302]))
    np.savez_compressed(os.path.join(HERE, "csf_res2net.npz"), __meta__=np.array(json.dumps(meta)), **out)
    print("wrote csf_res2net.npz", {k: v.tolist() for k, v in out.items() if k.endswith("stats")})


if __name__ == "__main__":
    if "--train-only" in sys.argv:
        main_train()
    else:
        if "--r-only" not in sys.argv:
            main()
            main_train()
        main_r()
