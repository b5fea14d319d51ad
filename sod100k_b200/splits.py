"""Channel-split arithmetic of the reference, reproduced with the same float operations.

Branch boundaries inside a gOctaveConv weight are `int(round(C * cumulative_alpha))` with python-float
running sums of alphas like 0.5816993464052287 (CSNet/model/csnet.py:641-650, 683-691); the alphas
themselves are `split * 1.0 / int(round(sum(split)))` (ILBlock.__init__, csnet.py:26-31).  Re-deriving
the cuts from the integer config instead can move a boundary by one channel, so follow the floats.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

DILATIONS = (1, 2, 4, 8, 16)    # MSBlock default, csnet.py:121


def total(split) -> int:
    return int(round(float(np.sum(np.asarray(split)))))


def alphas(split) -> List[float]:
    split = np.asarray(split)
    return (split * 1.0 / total(split)).tolist()


def cumulative(alpha: Sequence[float]) -> List[float]:
    """[0, a0, a0+a1, ...] as python-float running sums (gOctaveConv.__init__, csnet.py:641-650)."""
    acc, run = [0], 0
    for a in alpha:
        run += a
        acc.append(run)
    return acc


def cuts(channels: int, alpha: Sequence[float]) -> List[int]:
    return [int(round(channels * c)) for c in cumulative(alpha)]


def widths(channels: int, alpha: Sequence[float]) -> List[int]:
    """Per-branch BN/PReLU widths: int(round(C * alpha_i)), non-cumulative (csnet.py:762-767, 815-828)."""
    return [int(round(channels * a)) for a in alpha]


def block_walk(layer_config):
    """(prefix, config index, stride, first) for every ILBlock in CSNet.__init__ order (csnet.py:213-302),
    and the index of the first CSF-head entry."""
    stages = [int(s) for s in layer_config[-1]]
    walk, idx = [("stage0.0", 0, 1, True)], 1
    for s in range(4):
        for k in range(stages[s]):
            walk.append((f"stage{s + 1}.{k}", idx, 2 if (s > 0 and k == 0) else 1, False))
            idx += 1
    return walk, idx
