"""core/ml/SplitStrategy.scala -- how the master partitions row ids over workers."""
from __future__ import annotations

import math
from typing import List


def vanilla(n_rows: int, n_slaves: int) -> List[range]:
    """SplitStrategy.vanilla (SplitStrategy.scala:13-14): `indices.grouped(ceil(n / K))` -- contiguous
    groups; there may be FEWER than K groups and the last one may be short (vanilla(9, 4) has 3)."""
    size = int(math.ceil(n_rows / float(n_slaves)))
    return [range(s, min(s + size, n_rows)) for s in range(0, n_rows, size)]
