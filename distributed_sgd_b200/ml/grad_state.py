"""core/ml/GradState.scala -- what `fit` returns.  `grad` holds the WEIGHTS (the reference's field name
is kept, quirk Q8)."""
from __future__ import annotations

import time
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np


def _now_ms() -> int:
    return int(time.time() * 1000)


@dataclass(frozen=True)
class GradState:
    grad: np.ndarray
    loss: Optional[float] = None
    start: int = 0
    updates: int = 0
    end: Optional[int] = None

    @staticmethod
    def start_state(weights: np.ndarray) -> "GradState":      # GradState.start (GradState.scala:19-23)
        return GradState(weights, None, _now_ms(), 0, None)

    def replace_grad(self, new_weights: np.ndarray) -> "GradState":  # GradState.scala:10 (also bumps `updates`)
        return replace(self, grad=new_weights, updates=self.updates + 1)

    def finish(self, final_loss: float) -> "GradState":       # GradState.scala:12
        return replace(self, loss=final_loss, end=_now_ms())
