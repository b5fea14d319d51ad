"""Algorithmic (compulsory) HBM traffic of the CSNet forward, from the layer_config alone.

SURVEY.md §8(d): the primary figure is BLOCK-fused traffic — for each of the reference's fusion-closed
units (every ILBlock, the CSFHead, cls+final upsample) the distinct input elements read once plus the
output elements written once, un-padded channel counts; weights are amortised over the batch and ignored.
The secondary figure is MODULE-fused (gOctaveCBR / SimplifiedGOctConvBR / MSBlock / cls as the units).
csnet-L-x2 at 224x224: 15.734 M block elements, 42.414 M module elements per image.
"""
from __future__ import annotations

import numpy as np

from . import splits


def _branch_elems(split, H, W):
    """Σ_b C_b * (H/2^b) * (W/2^b) over live branches of a split."""
    return [int(round(float(c))) * (H >> b) * (W >> b) for b, c in enumerate(np.asarray(split).reshape(-1))]


def forward_elements(layer_config, H: int, W: int):
    """Returns dict(block=..., module=...) of per-image element counts."""
    walk, idx = splits.block_walk(layer_config)
    block = module = 0
    res = (H, W)                      # resolution of branch 0 of the current block INPUT
    feats = {}
    for prefix, ci, stride, first in walk:
        in_split = np.array([3]) if first else layer_config[ci][0]
        out_split = layer_config[ci][1]
        i_el = sum(_branch_elems(in_split, *res))
        if stride == 2:
            res = (res[0] // 2, res[1] // 2)
        o_el = sum(_branch_elems(out_split, *res))
        block += i_el + o_el
        module += (i_el + o_el) + 2 * (o_el + o_el)     # conv1x1 unit + two depthwise units
        feats[prefix] = (int(round(float(np.asarray(out_split).reshape(-1)[0]))), res)
    stages = [int(s) for s in layer_config[-1]]
    ends = [f"stage{s + 1}.{stages[s] - 1}" for s in (1, 2, 3)]
    c_in, c_mid, c_out = layer_config[idx:idx + 3]
    head_in = sum(c * r[0] * r[1] for c, r in (feats[e] for e in ends))
    rs = [feats[e][1] for e in ends]
    mid_in = sum(int(c) * r[0] * r[1] for c, r in zip(np.asarray(c_mid[0]).reshape(-1), rs))
    mid_out = sum(int(c) * r[0] * r[1] for c, r in zip(np.asarray(c_mid[1]).reshape(-1), rs))
    fuse_out = splits.total(c_out[1]) * rs[0][0] * rs[0][1]
    block += head_in + fuse_out
    module += (head_in + mid_in) + (mid_in + mid_out) + (mid_out + fuse_out)
    tail = fuse_out + H * W
    block += tail
    module += tail
    return dict(block=block, module=module)


def bytes_per_image(layer_config, H, W, dtype_bytes: int, unit: str = "block") -> int:
    return forward_elements(layer_config, H, W)[unit] * dtype_bytes
