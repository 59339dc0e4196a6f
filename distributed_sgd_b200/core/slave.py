"""core/Slave.scala -- one worker.  Here: one GPU, one dsgd_ctx.

The reference Slave owns the whole training array and answers `forward` / `gradient` /
`startAsync` / `updateGrad` / `stopAsync` RPCs (core/Slave.scala:113-197).  This class keeps that
surface (same method names and argument meaning, snake_case) and hands every one of them to the CUDA
library through the C ABI; there is no arithmetic in this file.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from ..ml.sparse_svm import SparseSVM
from ..native import NativeCtx
from ..utils.dataset import Data


class Slave:
    def __init__(self, node: int, master: int, data: Data, model: SparseSVM, is_async: bool = False, *,
                 world: int = 1, device: Optional[int] = None, test_data: Optional[Data] = None,
                 ctx: Optional[NativeCtx] = None):
        """`new Slave(node, master, data, model, async)` (core/Slave.scala:20; Main.scala:138,149).

        node = this worker's rank; master = the master's rank (kept for recognisability; the master logic
        runs SPMD on every rank).  `data` is the FULL training array, addressed by global row id (quirk
        Q13).  test_data (extension): rows appended after the training rows so the same device context can
        serve Master.localLoss(testData) -- they are never sampled.  ctx (extension): a device context that already
        holds exactly these rows (train rows followed by the test rows) and its dimSparsity.
        """
        self.node, self.master, self.model, self.is_async, self.world = node, master, model, is_async, world
        self.n_train = data.n_rows
        self.n_test = test_data.n_rows if test_data is not None else 0
        self.dim = data.dim
        if ctx is not None:
            if ctx.n_rows != self.n_train + self.n_test or ctx.dim != data.dim:
                raise ValueError("Slave: the given device context does not hold these rows")
            self.ctx = ctx
            if model.dim_sparsity is None:
                model.dim_sparsity = ctx.compute_dim_sparsity(self.n_train)
            return
        self.ctx = NativeCtx(node if device is None else device, data.dim, model.lam, rank=node, world=world,
                             is_async=is_async)
        if test_data is not None:
            row_ptr = np.concatenate([data.row_ptr, test_data.row_ptr[1:] + data.row_ptr[-1]])
            col = np.concatenate([data.col[:data.nnz], test_data.col[:test_data.nnz]])
            val = np.concatenate([data.val[:data.nnz], test_data.val[:test_data.nnz]])
            label = np.concatenate([data.label, test_data.label])
            self.ctx.load_csr(row_ptr, col, val, label)
        else:
            self.ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
        if model.dim_sparsity is None:
            model.dim_sparsity = self.ctx.compute_dim_sparsity(self.n_train)  # Main.scala:54-65 on the device
        else:
            self.ctx.set_dim_sparsity(model.dim_sparsity)

    def stop(self):  # Slave.stop (core/Slave.scala:68-77): releases the device context
        self.ctx.close()

    # ---- SlaveImpl -----------------------------------------------------------------------------------
    def forward(self, samples_idx: Sequence[int], weights: Optional[np.ndarray] = None) -> np.ndarray:
        """SlaveImpl.forward (core/Slave.scala:129-140): predictions -signum(x.w) for the listed rows."""
        self._train_ids(samples_idx)
        return self.ctx.forward(samples_idx, weights)

    def gradient(self, weights: Optional[np.ndarray], samples_idx: Sequence[int]) -> np.ndarray:
        """SlaveImpl.gradient (core/Slave.scala:142-157): regularize(sum of backward over the batch)."""
        self._train_ids(samples_idx)
        return self.ctx.gradient(samples_idx, weights)

    def start_async(self, weights: np.ndarray, samples: Sequence[int], batch_size: int, learning_rate: float, *,
                    concurrency: int = 1, max_updates: int = 0, seed: int = 0):
        """SlaveImpl.startAsync (core/Slave.scala:159-175)."""
        self._train_ids(samples)
        self.ctx.start_async(weights, samples, batch_size, learning_rate, concurrency, max_updates, seed)

    def update_grad(self, grad_update):
        """SlaveImpl.updateGrad (core/Slave.scala:177-185): weights -= gradUpdate.  Accepts a dense vector or
        an (indices, values) pair."""
        if isinstance(grad_update, tuple):
            idx, val = grad_update
        else:
            dense = np.asarray(grad_update, dtype=np.float64)
            idx = np.flatnonzero(dense).astype(np.int32)
            val = dense[idx]
        self.ctx.update_grad(idx, val)

    def stop_async(self):
        """SlaveImpl.stopAsync (core/Slave.scala:187-195)."""
        self.ctx.stop_async()

    # ---- helpers ---------------------------------------------------------------------------------------
    def _train_ids(self, idx):
        idx = np.asarray(idx)
        if idx.size and (idx.min() < 0 or idx.max() >= self.n_train):
            raise IndexError("sample id outside the training rows")  # data(idx) on the reference's array
