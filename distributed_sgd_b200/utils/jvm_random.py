"""JVM-exact batch draws: java.util.Random + scala.util.Random.shuffle (Scala 2.12), so that a run
here can draw sample-for-sample the batches a reference run draws after `Random.setSeed(0)` (Main.scala:32;
core/Master.scala:184-187).  The generator is pinned on java.util.Random's well-known outputs
(tests/test_host_logic.py); the shuffle order follows the Scala 2.12 source from memory and cannot be cross-checked in this
image (no JVM) -- treat trajectories obtained with it as "expected to match", not "verified to match"."""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from .. import native


class JvmRandom:
    """java.util.Random(seed): 48-bit LCG; `next_int()` / `next_int(bound)` like the Java methods."""

    def __init__(self, seed: int = 0):
        self._h = native.host_lib()
        self._h.dsgd_jrandom_next_int.restype = C.c_int32
        self._h.dsgd_jvm_sync_epoch.restype = C.c_int64
        self._state = C.c_uint64()
        self._h.dsgd_jrandom_seed(C.byref(self._state), C.c_int64(seed))

    def next_int(self, bound: int = 0) -> int:
        return int(self._h.dsgd_jrandom_next_int(C.byref(self._state), C.c_int32(bound)))

    def shuffle(self, xs) -> np.ndarray:
        """scala.util.Random.shuffle(xs): a new shuffled copy."""
        buf = np.ascontiguousarray(xs, dtype=np.int32).copy()
        self._h.dsgd_scala_shuffle_i32(C.byref(self._state), buf.ctypes.data_as(C.c_void_p), C.c_int64(buf.size))
        return buf

    def sync_epoch(self, n_rows: int, n_slaves: int, batch_size: int, group_size: int = 0) -> List[List[np.ndarray]]:
        """The draws of one epoch of Master.fit over SplitStrategy.vanilla(n_rows, n_slaves): list of steps, each a list
        of per-group arrays (shorter or empty at the tail, like `idx.slice(batch, batch + batchSize)`)."""
        group = group_size or -(-n_rows // n_slaves)       # ceil(n / K)  (core/ml/SplitStrategy.scala:14)
        n_groups = -(-n_rows // group)
        steps = -(-min(group, n_rows) // batch_size)
        out = np.empty(steps * n_groups * batch_size, dtype=np.int32)
        got = self._h.dsgd_jvm_sync_epoch(C.byref(self._state), C.c_int64(n_rows), C.c_int64(group), C.c_int32(batch_size),
                                          out.ctypes.data_as(C.c_void_p), C.c_int64(out.size))
        if got != steps:
            raise RuntimeError(f"dsgd_jvm_sync_epoch failed ({got})")
        out = out.reshape(steps, n_groups, batch_size)
        return [[out[s, k][out[s, k] >= 0].copy() for k in range(n_groups)] for s in range(steps)]

    def async_draws(self, assigned, n_updates: int, batch_size: int = 1) -> np.ndarray:
        """The row ids n_updates iterations of Slave.asyncTask draw (core/Slave.scala:83-88): batch 1 is
        `assignedSamples(Random.nextInt(size))`; batch > 1 is `Random.shuffle(assignedSamples.indices) take batchSize`, i.e.
        POSITIONS used as row ids (quirk Q6).  Feed the result to `NativeCtx.async_replay` (one lane, deterministic)."""
        assigned = np.ascontiguousarray(assigned, dtype=np.int32)
        if batch_size == 1:
            return np.array([assigned[self.next_int(assigned.size)] for _ in range(n_updates)], dtype=np.int32)
        pos = np.arange(assigned.size, dtype=np.int32)
        return np.concatenate([self.shuffle(pos)[:batch_size] for _ in range(n_updates)]).astype(np.int32)
