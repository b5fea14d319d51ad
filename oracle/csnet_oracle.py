"""TEST INFRASTRUCTURE — CPU restatement of the reference CSNet forward / train step.

This file is the parity ORACLE.  It is not part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
The product (`sod100k_b200/`) never does, and fails loudly when its CUDA library is missing.

What it restates: the function `x[N,3,H,W] -> logits[N,1,H,W]` of
`/root/reference/CSNet/model/csnet.py` (CSNet.forward :365-387) written as one flat functional
program over a plain `{state_dict key: tensor}` mapping and the `layer_config` list, in fp32 on the
CPU.  All arithmetic of the reference lives in torch ATen calls (F.conv2d, F.batch_norm, F.prelu,
F.avg_pool2d, F.max_pool2d, F.interpolate — SURVEY.md §2.1 K1-K9); the reference pins no torch
version and holds no golden vectors of its own, so the oracle is *defined* as "the same ATen calls
on torch 2.11 CPU fp32" and is pinned against outputs of the unmodified reference run in the build
container (tests/golden/*.npz, produced by tests/golden/make_golden.py).

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d default, created at csnet.py:138,764,825
BN_MOMENTUM = 0.1      # nn.BatchNorm2d default
DILATIONS = (1, 2, 4, 8, 16)   # csnet.py:121


# --------------------------------------------------------------------------------------------------
# channel-split arithmetic
# --------------------------------------------------------------------------------------------------
def alphas_of(split) -> List[float]:
    """ILBlock.__init__ (csnet.py:26-31): alpha = split * 1.0 / int(round(sum(split))), as a list."""
    split = np.asarray(split)
    total = int(round(float(np.sum(split))))
    return (split * 1.0 / total).tolist()


def total_of(split) -> int:
    return int(round(float(np.sum(np.asarray(split)))))


def cumulative(alphas: Sequence[float]) -> List[float]:
    """gOctaveConv.__init__ (csnet.py:641-650): running python-float sums [0, a0, a0+a1, ...]."""
    out, s = [0], 0
    for a in alphas:
        s += a
        out.append(s)
    return out


def channel_cuts(channels: int, alphas: Sequence[float]) -> List[int]:
    """gOctaveConv.forward (csnet.py:683-691): boundaries int(round(C * cumulative_alpha))."""
    return [int(round(channels * c)) for c in cumulative(alphas)]


# --------------------------------------------------------------------------------------------------
# primitive modules
# --------------------------------------------------------------------------------------------------
class _Ctx:
    """Carries mode flags, the parameter mapping and optional recorders through the program."""

    def __init__(self, sd, training, taps, new_stats, reg):
        self.sd = sd
        self.training = training
        self.taps = taps
        self.new_stats = new_stats
        self.reg = reg          # None or dict(expandflop=..., terms=[...])

    def tap(self, name, value):
        if self.taps is not None:
            self.taps[name] = value


def _bn_prelu(ctx: _Ctx, x: torch.Tensor, bn: str, prelu: str) -> torch.Tensor:
    """BatchNorm2d then PReLU (csnet.py:786,791,846-847,148).

    eval: running statistics; train: batch statistics, and the momentum-0.1 / unbiased-variance
    update of the running statistics is reported in ctx.new_stats (the mapping itself is not mutated).
    """
    sd = ctx.sd
    rm, rv = sd[bn + ".running_mean"], sd[bn + ".running_var"]
    if ctx.training:
        rm, rv = rm.detach().clone(), rv.detach().clone()
        y = F.batch_norm(x, rm, rv, sd[bn + ".weight"], sd[bn + ".bias"], True, BN_MOMENTUM, BN_EPS)
        if ctx.new_stats is not None:
            ctx.new_stats[bn + ".running_mean"] = rm
            ctx.new_stats[bn + ".running_var"] = rv
    else:
        y = F.batch_norm(x, rm, rv, sd[bn + ".weight"], sd[bn + ".bias"], False, BN_MOMENTUM, BN_EPS)
    return F.prelu(y, sd[prelu + ".weight"])


def goct_conv(xs: List[Optional[torch.Tensor]], w: torch.Tensor, alpha_in, alpha_out,
              pad: int, stride: int) -> List[Optional[torch.Tensor]]:
    """gOctaveConv.forward (csnet.py:664-726).

    in-branch i lives at resolution H/2^i; out-branch j at H/2^j (after the optional stride-2
    average pool that is applied to EVERY input branch, :679-680).  Path (i, j):
      i > j  conv at the low resolution, then bilinear x2^(i-j)          (:702-707)
      i < j  max-pool 2^(j-i) first, then conv                          (:708-714)
      i == j conv                                                         (:715-717)
    and out_j is the python `sum` of its paths in increasing i (:720-724).
    """
    cout, cin = int(w.shape[0]), int(w.shape[1])
    ci, co = channel_cuts(cin, alpha_in), channel_cuts(cout, alpha_out)
    outs: List[Optional[torch.Tensor]] = [None] * len(alpha_out)
    for i, xi in enumerate(xs):
        if xi is None:
            continue
        if stride == 2:
            xi = F.avg_pool2d(xi, (2, 2), stride=2)
        if ci[i] == ci[i + 1]:
            continue
        for j in range(len(alpha_out)):
            if co[j] == co[j + 1]:
                continue
            wij = w[co[j]:co[j + 1], ci[i]:ci[i + 1]]
            if i > j:
                y = F.conv2d(xi, wij, None, 1, pad)
                y = F.interpolate(y, scale_factor=2 ** (i - j), mode="bilinear")
            elif i < j:
                k = 2 ** (j - i)
                y = F.conv2d(F.max_pool2d(xi, k, stride=k), wij, None, 1, pad)
            else:
                y = F.conv2d(xi, wij, None, 1, pad)
            outs[j] = (0 + y) if outs[j] is None else outs[j] + y
    return outs


def goct_cbr(ctx: _Ctx, prefix: str, xs, alpha_in, alpha_out, ksize: int, stride: int):
    """gOctaveCBR.forward (csnet.py:778-792) — conv, then per-branch BN + PReLU.

    A single-in / single-out block is a plain `Conv2dX100` (csnet.py:751-754): weight x100
    (conv2d.py:104) and a real strided convolution instead of the average pool.
    """
    pad = 1 if ksize == 3 else 0
    w = ctx.sd[prefix + ".conv.weight"]
    if len(alpha_in) == 1 and len(alpha_out) == 1:
        y = F.conv2d(xs[0], 100.0 * w, None, stride, pad)
        outs = [_bn_prelu(ctx, y, prefix + ".bns.0", prefix + ".prelus.0")]
    else:
        outs = goct_conv(xs, w, alpha_in, alpha_out, pad, stride)
        for j, y in enumerate(outs):
            if y is not None:
                outs[j] = _bn_prelu(ctx, y, f"{prefix}.bns.{j}", f"{prefix}.prelus.{j}")
    _flops_term(ctx, prefix, outs)
    return outs


def dw_cbr(ctx: _Ctx, prefix: str, xs):
    """SimplifiedGOctConvBR.forward (csnet.py:838-851): per-branch depthwise 3x3 (Conv2dX100,
    groups=C, pad 1, weight x100) -> BN -> PReLU."""
    outs = []
    for b, x in enumerate(xs):
        if x is None:
            outs.append(None)
            continue
        w = ctx.sd[f"{prefix}.convs.{b}.weight"]
        y = F.conv2d(x, 100.0 * w, None, 1, 1, 1, int(w.shape[0]))
        outs.append(_bn_prelu(ctx, y, f"{prefix}.bns.{b}", f"{prefix}.prelus.{b}"))
    _flops_term(ctx, prefix, outs)
    return outs


def _flops_term(ctx: _Ctx, prefix: str, outs):
    """Oct_bn_hook (csnet.py:391-410), registered on ILBlock.{conv1x1,conv3x3_1,conv3x3_2} only
    (flops_hook :332-355).  term = 0.5 * sum_b w_b * sum(|GAP(out_b.detach())| * gamma_b^2) with
    w_b = baseflop * expandflop^(branches-1-b).  For N>1 `squeeze()` leaves [N,C] which broadcasts
    against gamma[C], i.e. the term is summed over the batch (get_flops divides by batchsize, :330).
    """
    reg = ctx.reg
    if reg is None or not prefix.startswith("stage"):
        return
    base = reg["baseflop"][prefix.rsplit(".", 1)[0]]
    e = reg["expandflop"]
    branches = len(outs)
    wts, f = [], base * (e ** (branches - 1))
    for _ in range(branches):
        wts.append(f)
        f /= e
    terms = []
    for b, y in enumerate(outs):
        if y is None:
            continue
        gap = F.adaptive_avg_pool2d(y.detach(), 1).squeeze().abs()
        gamma = ctx.sd[f"{prefix}.bns.{b}.weight"]
        terms.append((wts[b] * gap * torch.pow(gamma, 2)).sum())
    reg["terms"].append(0.5 * sum(terms))


def il_block(ctx: _Ctx, prefix: str, xs, in_split, out_split, stride: int, first: bool):
    """ILBlock (csnet.py:17-76): gOctaveCBR (3x3 pad 1 when `first` or stride 2, else 1x1) followed by
    two depthwise 3x3 conv+BN+PReLU layers.  No residual."""
    a_in, a_out = alphas_of(in_split), alphas_of(out_split)
    k = 3 if (first or stride == 2) else 1
    y = goct_cbr(ctx, prefix + ".conv1x1", xs, a_in, a_out, k, stride)
    y = dw_cbr(ctx, prefix + ".conv3x3_1", y)
    y = dw_cbr(ctx, prefix + ".conv3x3_2", y)
    ctx.tap(prefix, y)
    return y


def ms_block(ctx: _Ctx, prefix: str, x, dil_channels):
    """MSBlock.forward (csnet.py:141-149): parallel dilated 3x3 convs (Conv2dX100, pad = dilation,
    entries pruned to 0 channels are skipped, :127-137), channel concat, BN, PReLU."""
    outs = []
    for d, dil in enumerate(DILATIONS):
        if int(dil_channels[d]) != 0:
            w = ctx.sd[f"{prefix}.msconv.{d}.weight"]
            outs.append(F.conv2d(x, 100.0 * w, None, 1, dil, dil))
    return _bn_prelu(ctx, torch.cat(outs, dim=1), prefix + ".bn", prefix + ".prelu")


def csf_head(ctx: _Ctx, prefix: str, xs, cfg3):
    """CSFHead (csnet.py:153-206): fuse (3->3 gOctaveCBR 1x1) -> PallMSBlock (:79-113) -> fuse1x1
    (3->1 gOctaveCBR 1x1)."""
    a_in = alphas_of(cfg3[0][0])
    a_mid_in = alphas_of(cfg3[1][0])
    a_mid_out = alphas_of(cfg3[1][1])
    dils = np.asarray(cfg3[1][2])
    y = goct_cbr(ctx, prefix + ".fuse", xs, a_in, a_mid_in, 1, 1)
    ctx.tap(prefix + ".fuse", y)
    z = []
    for b in range(len(a_mid_in)):
        if max(dils[b]) != 0:
            z.append(ms_block(ctx, f"{prefix}.ms.convs.{b}", y[b], dils[b]))
        else:
            z.append(None)
    ctx.tap(prefix + ".ms", z)
    out = goct_cbr(ctx, prefix + ".fuse1x1", z, a_mid_out, [1], 1, 1)
    ctx.tap(prefix + ".fuse1x1", out)
    return out


def block_table(layer_config):
    """The (prefix, config index, stride, first) walk of CSNet.__init__ (csnet.py:213-302)."""
    stages = list(layer_config[-1])
    table, idx = [("stage0.0", 0, 1, True)], 1
    for s in range(4):
        for k in range(stages[s]):
            table.append((f"stage{s + 1}.{k}", idx, 2 if (s > 0 and k == 0) else 1, False))
            idx += 1
    return table, idx


def flops_baseflops(layer_config, expandflop: float) -> Dict[str, float]:
    """CSNet.flops_hook (csnet.py:332-355): per-ILBlock `baseflop`, starting at
    expandflop^(len(stages)-1) and divided by expandflop after stage0+stage1, stage2, stage3 ..."""
    stages = list(layer_config[-1])
    real = stages.copy()
    real[0] += 1
    table, _ = block_table(layer_config)
    base, out, stage, in_stage = expandflop ** (len(stages) - 1), {}, 0, 0
    for prefix, *_ in table:
        out[prefix] = base
        in_stage += 1
        if in_stage == real[stage]:
            base /= expandflop
            stage += 1
            in_stage = 0
    return out


def csnet_forward(layer_config, sd: Dict[str, torch.Tensor], x: torch.Tensor, *, training: bool = False,
                  taps: Optional[dict] = None, new_stats: Optional[dict] = None,
                  flops_expand: Optional[float] = None):
    """CSNet.forward (csnet.py:365-387).  Returns logits [N,1,H,W]; with `flops_expand` set also
    returns the un-normalised regulariser sum (what `get_flops()*batchsize` would be, :324-330)."""
    reg = None
    if flops_expand is not None:
        reg = dict(expandflop=flops_expand, baseflop=flops_baseflops(layer_config, flops_expand), terms=[])
    ctx = _Ctx(sd, training, taps, new_stats, reg)
    table, idx = block_table(layer_config)
    feats, cur = {}, [x]
    for prefix, ci, stride, first in table:
        in_split = np.array([3]) if first else layer_config[ci][0]
        cur = il_block(ctx, prefix, cur, in_split, layer_config[ci][1], stride, first)
        feats[prefix] = cur
    stages = list(layer_config[-1])
    ends = [f"stage{s + 1}.{stages[s] - 1}" for s in (1, 2, 3)]
    fuse = csf_head(ctx, "oct_fuse", [feats[e][0] for e in ends], layer_config[idx:idx + 3])
    out = F.conv2d(fuse[0], sd["cls_layer.weight"], sd["cls_layer.bias"])          # :306-308,381
    out = F.interpolate(out, x.shape[2:], mode="bilinear", align_corners=False)    # :382-385
    ctx.tap("logits", out)
    if reg is not None:
        return out, sum(reg["terms"])
    return out


# --------------------------------------------------------------------------------------------------
# state_dict shapes (so tests can synthesise parameters for any layer_config without the reference)
# --------------------------------------------------------------------------------------------------
def state_shapes(layer_config) -> Dict[str, tuple]:
    """Key -> shape of CSNet(layer_config).state_dict() (module constructors
    csnet.py:18-70,116-139,153-200,733-776,799-836)."""
    shapes: Dict[str, tuple] = {}

    def bn(prefix, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{prefix}.{k}"] = (c,)
        shapes[f"{prefix}.num_batches_tracked"] = ()

    def cbr(prefix, cin, cout, a_out, k):
        shapes[prefix + ".conv.weight"] = (cout, cin, k, k)
        for j, a in enumerate(a_out):
            c = int(round(cout * a))
            if c != 0:
                bn(f"{prefix}.bns.{j}", c)
                shapes[f"{prefix}.prelus.{j}.weight"] = (c,)

    def dw(prefix, cout, a_out):
        for j, a in enumerate(a_out):
            c = int(round(cout * a))
            if c >= 1:
                shapes[f"{prefix}.convs.{j}.weight"] = (c, 1, 3, 3)
                bn(f"{prefix}.bns.{j}", c)
                shapes[f"{prefix}.prelus.{j}.weight"] = (c,)

    table, idx = block_table(layer_config)
    for prefix, ci, stride, first in table:
        in_split = np.array([3]) if first else layer_config[ci][0]
        cin, cout = total_of(in_split), total_of(layer_config[ci][1])
        a_out = alphas_of(layer_config[ci][1])
        cbr(prefix + ".conv1x1", cin, cout, a_out, 3 if (first or stride == 2) else 1)
        dw(prefix + ".conv3x3_1", cout, a_out)
        dw(prefix + ".conv3x3_2", cout, a_out)
    c0, c1, c2 = layer_config[idx:idx + 3]
    fin, fmid, fmo, fout = total_of(c0[0]), total_of(c1[0]), total_of(c1[1]), total_of(c2[1])
    cbr("oct_fuse.fuse", fin, fmid, alphas_of(c1[0]), 1)
    a_mi, a_mo, dils = alphas_of(c1[0]), alphas_of(c1[1]), np.asarray(c1[2])
    for b in range(len(a_mi)):
        if max(dils[b]) != 0:
            cin_b, cout_b = int(round(fmid * a_mi[b])), int(round(fmo * a_mo[b]))
            for d in range(len(DILATIONS)):
                if int(dils[b][d]) != 0:
                    shapes[f"oct_fuse.ms.convs.{b}.msconv.{d}.weight"] = (int(dils[b][d]), cin_b, 3, 3)
            bn(f"oct_fuse.ms.convs.{b}.bn", cout_b)
            shapes[f"oct_fuse.ms.convs.{b}.prelu.weight"] = (cout_b,)
    cbr("oct_fuse.fuse1x1", fmo, fout, [1], 1)
    shapes["cls_layer.weight"] = (1, total_of(layer_config[-2][1]), 1, 1)
    shapes["cls_layer.bias"] = (1,)
    return shapes


def init_layer_config(basewidth: int, basic_split):
    """init_layers (csnet.py:414-518): the un-pruned layer_config for a given base width/split."""
    bs = np.array([float(v) for v in basic_split])
    one = np.array([1.0])
    stages = [3, 4, 6, 4]
    cfg = [[np.array([3]), basewidth * bs]]
    cfg += [[basewidth * bs, basewidth * bs] for _ in range(stages[0])]
    for mult, n in ((2, stages[1]), (4, stages[2])):
        prev = mult // 2
        cfg.append([basewidth * prev * (one if mult == 4 else bs), basewidth * mult * bs])
        cfg += [[basewidth * mult * bs, basewidth * mult * bs] for _ in range(1, n - 1)]
        cfg.append([basewidth * mult * bs, basewidth * mult * one])
    cfg.append([basewidth * 4 * one, basewidth * 4 * bs])                     # stage 4
    cfg += [[basewidth * 4 * bs, basewidth * 4 * bs] for _ in range(1, stages[3] - 1)]
    cfg.append([basewidth * 4 * bs, basewidth * 4 * one])
    s2, s3, s4 = basewidth * 2, basewidth * 4, basewidth * 4
    mid = np.array([s2 // 3, s3 // 3, s4 // 3])
    dil = []
    for br in mid:
        each = br // len(DILATIONS)
        dil.append([each] * (len(DILATIONS) - 1) + [br - each * (len(DILATIONS) - 1)])
    cfg.append([np.array([s2, s3, s4]), mid.copy()])
    cfg.append([mid.copy(), mid.copy(), np.array(dil)])
    cfg.append([mid.copy(), np.array([int(mid.sum())])])
    for c in cfg:
        c[0] = np.round(c[0]).astype(np.int32)
        c[1] = np.round(c[1]).astype(np.int32)
    cfg.append(stages)
    return cfg


# --------------------------------------------------------------------------------------------------
# train step (T/train.py:97-123, 203-216)
# --------------------------------------------------------------------------------------------------
def zero_wd_names(names: Sequence[str]) -> List[str]:
    """T/train.py:101-105 — BN gammas of ILBlock.conv1x1 and conv3x3_1 get weight_decay 0.  The
    reference tests 'conv3x3_1.bns' twice and never 'conv3x3_2.bns'; that is reproduced."""
    return [n for n in names
            if "stage" in n and ("conv1x1.bns" in n or "conv3x3_1.bns" in n or "conv3x3_1.bns" in n) and "weight" in n]


def adam_update(p, g, m, v, step, lr, wd, beta1=0.9, beta2=0.99, eps=1e-8):
    """torch.optim.Adam single-tensor rule with L2 weight decay folded into the gradient (not AdamW):
    g += wd*p; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g^2;
    p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).   (T/train.py:108-123)"""
    g = g + wd * p if wd != 0 else g
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


def train_step(layer_config, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor],
               opt: dict, x: torch.Tensor, target: torch.Tensor, *, lr=1e-4, weight_decay=5e-3,
               flops_weight: Optional[float] = None, flops_expand: float = 1.0):
    """One reference training step (T/train.py:203-216): train-mode forward, mean BCE-with-logits,
    optional `WEIGHT * get_flops()` regulariser (:212-213), backward, Adam in two weight-decay groups.
    Returns (loss_without_reg, grads, new_params, new_buffers, new_opt)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    sd = dict(buffers)
    sd.update(leaves)
    new_stats: Dict[str, torch.Tensor] = {}
    if flops_weight is not None:
        out, reg = csnet_forward(layer_config, sd, x, training=True, new_stats=new_stats, flops_expand=flops_expand)
    else:
        out, reg = csnet_forward(layer_config, sd, x, training=True, new_stats=new_stats), None
    loss = F.binary_cross_entropy_with_logits(out, target)
    total = loss if reg is None else loss + flops_weight * (reg / x.shape[0])
    names = list(leaves)
    grads = dict(zip(names, torch.autograd.grad(total, [leaves[n] for n in names], allow_unused=True)))
    nowd = set(zero_wd_names(names))
    step = opt.get("step", 0) + 1
    new_params, new_opt = {}, {"step": step, "m": {}, "v": {}}
    for n in names:
        g = grads[n] if grads[n] is not None else torch.zeros_like(params[n])
        grads[n] = g
        m = opt.get("m", {}).get(n, torch.zeros_like(g))
        v = opt.get("v", {}).get(n, torch.zeros_like(g))
        p, m, v = adam_update(params[n].detach(), g, m, v, step, lr, 0.0 if n in nowd else weight_decay)
        new_params[n], new_opt["m"][n], new_opt["v"][n] = p, m, v
    new_buffers = dict(buffers)
    new_buffers.update(new_stats)
    for k in buffers:
        if k.endswith("num_batches_tracked"):
            new_buffers[k] = buffers[k] + 1
    return loss.detach(), grads, new_params, new_buffers, new_opt
