"""One reference training step (CSNet_training/train.py:203-216) on the engine's kernels:
train-mode forward -> mean BCE-with-logits (+ WEIGHT * get_flops()) -> backward -> [DP: one all-reduce of the single
flat gradient bucket] -> Adam in the reference's two weight-decay groups (train.py:97-123).

Data parallelism: one process per GPU, local BatchNorm statistics (the reference has no SyncBN), the
only collective is the all-reduce(sum)/world of ONE flat fp32 bucket holding every gradient (140 894 floats for
csnet-L-x2) — every `p.grad` is a view into that bucket, so there is no flatten / unflatten copy.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch

from . import train_ops as T


def reference_param_groups(model) -> Tuple[List[torch.nn.Parameter], List[torch.nn.Parameter]]:
    """(normal, zero-weight-decay) exactly as train.py:101-105 picks them — including its repeated
    'conv3x3_1.bns' test (conv3x3_2's BN gammas stay in the decayed group)."""
    normal, picked = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:                  # frozen parameters take no optimizer step (no zero grads, no weight decay)
            continue
        if "stage" in name and ("conv1x1.bns" in name or "conv3x3_1.bns" in name or "conv3x3_1.bns" in name) and "weight" in name:
            picked.append(p)
        else:
            normal.append(p)
    return normal, picked


class FlatGrads:
    """All gradients of `params` as views of one contiguous fp32 bucket."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        self.bucket = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.bucket[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.bucket.zero_()

    def intact(self) -> bool:
        """autograd must have accumulated in place (the views still alias the bucket)."""
        base = self.bucket.data_ptr()
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                return False
            off += p.numel()
        return True

    def all_reduce_mean(self, group=None):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=group)
            self.bucket.div_(dist.get_world_size(group))


class Trainer:
    def __init__(self, model, lr: float = 1e-4, weight_decay: float = 5e-3, betas=(0.9, 0.99), eps: float = 1e-8,
                 flops_weight: Optional[float] = None, flops_expand: float = 1.0, process_group=None):
        self.model = model
        self.flops_weight = flops_weight
        self.group = process_group
        if flops_weight is not None:
            model.flops_hook(expandflop=flops_expand)
        self.flat = FlatGrads(model.parameters())
        normal, picked = reference_param_groups(model)
        self.opt = T.FusedAdam([(normal, weight_decay), (picked, 0.0)], lr=lr, betas=betas, eps=eps)

    def step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Returns the BCE loss (without the regulariser), like `losses.update(loss.item())` at train.py:211 —
        as a device tensor: no host sync inside the step."""
        m = self.model
        m.train()
        self.flat.zero()
        if self.flops_weight is not None:
            m.clear_flops()
            m.set_batchsize(x.shape[0])
        out = m(x)
        loss = T.BceFn.apply(out, target)
        total = loss if self.flops_weight is None else loss + self.flops_weight * m.get_flops()
        total.backward()
        if not self.flat.intact():
            raise T.runtime.EngineError("gradient views were replaced; the flat bucket is stale")
        self.flat.all_reduce_mean(self.group)
        self.opt.step()
        return loss.detach()
