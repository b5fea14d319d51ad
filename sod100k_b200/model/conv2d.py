"""`Conv2dX100` — parameter container with the reference's name, constructor and state_dict layout
(CSNet/model/conv2d.py:28-105).  The stored weight is 1/100 of the effective one; the x100 is folded into
the packed weights by sod100k_b200/compiler.py, never applied per call.  Its own forward is a single-op
program on the engine (used only when a caller invokes the leaf directly)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn import init


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class Conv2dX100(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=False, padding_mode="zeros"):
        super().__init__()
        if in_channels % groups or out_channels % groups:
            raise ValueError("in_channels and out_channels must be divisible by groups")
        if padding_mode != "zeros":
            raise NotImplementedError("only zero padding is used by CSNet")
        self.in_channels, self.out_channels, self.groups = in_channels, out_channels, groups
        self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.padding_mode = padding_mode
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            init.uniform_(self.bias, -1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in))

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, groups={self.groups}, x100")

    def forward(self, x):
        try:
            from ..modular import conv2d_x100_forward
        except ImportError:
            from sod100k_b200.modular import conv2d_x100_forward
        return conv2d_x100_forward(self, x)
