"""core/ml/EarlyStopping.scala -- stopping criteria over a newest-first list of losses (host control
logic of Master.fit; a handful of scalars per epoch, not part of the device hot path)."""
from __future__ import annotations

import sys
from typing import Callable, Optional, Sequence

EarlyStopping = Callable[[Sequence[float]], bool]


def target(target_loss: float) -> EarlyStopping:
    """EarlyStopping.target (EarlyStopping.scala:11): stop once the newest loss is <= target."""
    def crit(losses: Sequence[float]) -> bool:
        return len(losses) > 0 and losses[0] <= target_loss
    return crit


def no_improvement(patience: int = 5, min_delta: float = 1e-3, min_steps: Optional[int] = None) -> EarlyStopping:
    """EarlyStopping.noImprovement (EarlyStopping.scala:13-46).

    Scans newest -> oldest keeping a running minimum that is replaced whenever a value is within
    |minDelta| above it (so ties drift towards OLDER entries); stops iff the arg-min is not the newest
    entry and lies at least `patience` entries back.
    """
    tol = abs(min_delta)

    def arg_min(losses: Sequence[float]) -> int:
        best, where = sys.float_info.max, -1
        for i, v in enumerate(losses):
            if v - best <= tol:
                best, where = v, i
        return where

    def crit(losses: Sequence[float]) -> bool:
        if len(losses) == 0:
            return False
        if min_steps is not None and min_steps < len(losses):
            return False
        where = arg_min(losses)
        return where != 0 and where >= patience

    return crit
