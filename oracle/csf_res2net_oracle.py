"""TEST INFRASTRUCTURE — CPU restatement of the reference CSF+Res2Net forward (config 5).

Restates `/root/reference/CSF+Res2Net/networks/csf_res2net.py` (CSFNet.forward :251-259, Res2Net.forward :167-183,
Bottle2neck.forward :74-107, MSBlock.forward :218-225) and `networks/gOctConv.py` (gOctaveConv.forward :60-114,
gOctaveCBR.forward :140-152) as flat functions over a `{state_dict key: tensor}` mapping, fp32 (or any float dtype).
Pinned by tests/golden/csf_res2net.npz (outputs of the unmodified reference with seeded synthetic weights).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from .csnet_oracle import channel_cuts

FUSE_IN = (256 + 512 + 1024 + 2048, [1 / 15, 2 / 15, 4 / 15, 8 / 15])      # csf_res2net.py:239-240
FUSE_OUT = (128 + 256 + 512 + 512, [1 / 11, 2 / 11, 4 / 11, 4 / 11])       # :241-242
LAYERS = (3, 4, 6, 3)
DILATIONS = (1, 2, 4, 8, 16)


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.1, 1e-5)


def bottle2neck(sd, p: str, x, stride: int, stage: bool, width: int, scale: int = 4):
    """Bottle2neck.forward (:74-107): 1x1 -> split into `scale` groups of `width`; groups 0..scale-2 go through 3x3
    convs, hierarchically added (`sp = sp + spx[i]`) except in 'stage' blocks; last group passes (avg-pooled 3x3 in stage
    blocks); concat -> 1x1 -> + residual (avg-pool + 1x1 + BN shortcut when present) -> ReLU."""
    out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    spx = torch.split(out, width, 1)
    outs = []
    for i in range(scale - 1):
        sp = spx[i] if (i == 0 or stage) else sp + spx[i]
        sp = F.relu(_bn(sd, f"{p}.bns.{i}", F.conv2d(sp, sd[f"{p}.convs.{i}.weight"], None, stride, 1)))
        outs.append(sp)
    outs.append(F.avg_pool2d(spx[scale - 1], 3, stride, 1) if stage else spx[scale - 1])
    out = _bn(sd, p + ".bn3", F.conv2d(torch.cat(outs, 1), sd[p + ".conv3.weight"]))
    if p + ".downsample.1.weight" in sd:
        r = F.avg_pool2d(x, stride, stride, ceil_mode=True, count_include_pad=False) if stride > 1 else x
        r = _bn(sd, p + ".downsample.2", F.conv2d(r, sd[p + ".downsample.1.weight"]))
    else:
        r = x
    return F.relu(out + r)


def res2net50(sd, x, prefix="base"):
    """Res2Net.forward (:167-183): v1b stem (3 convs), max-pool 3x3/2, four stages; returns the four stage outputs."""
    p = prefix
    x = F.relu(_bn(sd, p + ".conv1.1", F.conv2d(x, sd[p + ".conv1.0.weight"], None, 2, 1)))
    x = F.relu(_bn(sd, p + ".conv1.4", F.conv2d(x, sd[p + ".conv1.3.weight"], None, 1, 1)))
    x = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.6.weight"], None, 1, 1)))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), LAYERS)):
        width = int(math.floor(planes * (26 / 64.0)))
        for b in range(blocks):
            x = bottle2neck(sd, f"{p}.layer{li + 1}.{b}", x, (1 if li == 0 else 2) if b == 0 else 1, b == 0, width)
        feats.append(x)
    return feats


def goct_conv_r(xs, w, alpha_in, alpha_out):
    """R/ gOctaveConv.forward (gOctConv.py:60-114), 1x1, stride 1: down paths resize the INPUT bilinearly to the
    output branch's size (:101-103), up paths resize the conv OUTPUT (:98-100)."""
    ci, co = channel_cuts(w.shape[1], alpha_in), channel_cuts(w.shape[0], alpha_out)
    outs = [None] * len(alpha_out)
    for i, x in enumerate(xs):
        if x is None or ci[i] == ci[i + 1]:
            continue
        for j in range(len(alpha_out)):
            if co[j] == co[j + 1]:
                continue
            wij = w[co[j]:co[j + 1], ci[i]:ci[i + 1]]
            size = xs[j].shape[2:4]
            if i > j:
                y = F.interpolate(F.conv2d(x, wij), size=size, mode="bilinear")
            elif i < j:
                y = F.conv2d(F.interpolate(x, size=size, mode="bilinear"), wij)
            else:
                y = F.conv2d(x, wij)
            outs[j] = (0 + y) if outs[j] is None else outs[j] + y
    return outs


def _gn_prelu(sd, gn, prelu, x):
    return F.prelu(F.group_norm(x, 32, sd[gn + ".weight"], sd[gn + ".bias"], 1e-5), sd[prelu + ".weight"])


def csf_head_r(sd, feats, taps=None):
    """CSFNet.forward after the backbone (:253-258)."""
    y = goct_conv_r(feats, sd["fuse.conv.weights"], FUSE_IN[1], FUSE_OUT[1])
    y = [_gn_prelu(sd, f"fuse.bns.{j}", f"fuse.prelus.{j}", t) for j, t in enumerate(y)]
    z = []
    for b, t in enumerate(y):                                   # PallMSBlock / MSBlock (:190-225)
        outs = [F.conv2d(t, sd[f"ms.convs.{b}.msconv.{d}.weight"], None, 1, dil, dil) for d, dil in enumerate(DILATIONS)]
        z.append(_gn_prelu(sd, f"ms.convs.{b}.bn", f"ms.convs.{b}.prelu", torch.cat(outs, 1)))
    f = goct_conv_r(z, sd["fuse1x1.conv.weights"], FUSE_OUT[1], [1])
    f0 = _gn_prelu(sd, "fuse1x1.bns.0", "fuse1x1.prelus.0", f[0])
    if taps is not None:
        taps.update(fuse=y, ms=z, fuse1x1=f0)
    return F.conv2d(f0, sd["cls_layer.weight"], sd["cls_layer.bias"])


def csfnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, taps=None):
    feats = res2net50(sd, x)
    if taps is not None:
        taps["feats"] = feats
    out = csf_head_r(sd, feats, taps)
    return F.interpolate(out, x.shape[2:], mode="bilinear", align_corners=False)
