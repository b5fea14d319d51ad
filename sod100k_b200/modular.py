"""Module-granular execution (training graph, callers that hook sub-modules).  Not built yet: fail loudly
rather than fall back to torch.nn.functional."""
from __future__ import annotations


def _nyi(what):
    raise NotImplementedError(f"{what}: module-granular / training execution on the B200 engine is not built yet "
                              "(inference runs through CSNet.forward in eval mode with no sub-module hooks)")


def csnet_forward(model, x):
    _nyi("CSNet.forward(train or hooked)")


def conv2d_x100_forward(m, x):
    _nyi("Conv2dX100.forward")


def goct_conv_forward(m, xset):
    _nyi("gOctaveConv.forward")


def goct_cbr_forward(m, xset):
    _nyi("gOctaveCBR.forward")


def dw_cbr_forward(m, xset):
    _nyi("SimplifiedGOctConvBR.forward")


def ms_block_forward(m, x):
    _nyi("MSBlock.forward")
