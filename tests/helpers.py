"""Shared helpers for the GPU parity tests (oracle side lives in oracle/, test infrastructure only)."""
import numpy as np

from oracle.oracle import Oracle


def make_pair(data, lam, n_train=None, device=0, rank=0, world=1, is_async=False):
    """(NativeCtx with `data` loaded and dimSparsity installed, Oracle with the same)."""
    from distributed_sgd_b200.native import NativeCtx
    n_train = data.n_rows if n_train is None else n_train
    ctx = NativeCtx(device, data.dim, lam, rank=rank, world=world, is_async=is_async)
    ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
    orc = Oracle(data.row_ptr, data.col, data.val, data.label, data.dim, lam)
    d = orc.dim_sparsity(n_train)
    orc.set_dim_sparsity(d)
    ctx.set_dim_sparsity(d)
    return ctx, orc


def data_from_csr(rp, col, val, lab, dim):
    from distributed_sgd_b200.utils.dataset import Data
    return Data(np.asarray(rp, np.int64), np.asarray(col, np.int32), np.asarray(val, np.float32),
                np.asarray(lab, np.int8), dim)
