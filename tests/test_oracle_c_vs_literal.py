"""Validates the fast array-based C oracle against the literal map-based restatement on random small
problems (including exact cancellations, gated-out batches, empty rows and multi-worker steps)."""
import numpy as np
import pytest

from oracle import scala_semantics as S
from oracle.oracle import Oracle, OracleError
from conftest import random_csr


def to_literal(row_ptr, col, val, label, dim):
    """CSR column c <-> reference key c+1 (RCV1 ids are 1-based; key == size is legal, quirk Q11)."""
    data = []
    for r in range(len(row_ptr) - 1):
        m = {int(col[p]) + 1: float(val[p]) for p in range(row_ptr[r], row_ptr[r + 1])}
        data.append((S.Sparse(m, dim), int(label[r])))
    return data


def w_to_literal(w, dim):
    return S.Sparse({j + 1: float(v) for j, v in enumerate(w)}, dim)


def literal_to_w(vec, dim):
    out = np.zeros(dim)
    for k, v in vec.map.items():
        out[k - 1] = v
    return out


@pytest.mark.parametrize("seed,dup", [(0, False), (1, True), (2, True), (3, False)])
def test_gradient_forward_loss(seed, dup):
    rng = np.random.default_rng(seed)
    dim, n = 40, 60
    rp, col, val, lab = random_csr(rng, n, dim, max_nnz=8, allow_empty=(seed == 3), dup_values=dup)
    lam = 0.05
    orc = Oracle(rp, col, val, lab, dim, lam)
    data = to_literal(rp, col, val, lab, dim)
    n_train = 45
    d_lit = S.dim_sparsity(data[:n_train])
    model = S.SparseSVM(lam, d_lit)
    d = orc.dim_sparsity(n_train)
    orc.set_dim_sparsity(d)

    for trial in range(6):
        w = np.where(rng.random(dim) < 0.6, rng.standard_normal(dim), 0.0) if trial else np.zeros(dim)
        wl = w_to_literal(w, dim)
        # the quirk-Q3 shift: literal w.dot(d) must equal the array dot in weight space
        assert np.isclose(wl.dot(d_lit), float(np.dot(w, d)), rtol=1e-13, atol=1e-300)
        idx = rng.choice(n, size=int(rng.integers(1, 25)), replace=False).astype(np.int32)
        g_ref = literal_to_w(S.slave_gradient(model, data, wl, idx.tolist()), dim)
        g, c = orc.gradient(w, idx)
        np.testing.assert_allclose(g, g_ref, rtol=1e-13, atol=0)
        assert (g == 0).tolist() == (g_ref == 0).tolist()          # identical support
        assert np.isclose(c, lam * 2.0 * wl.dot(d_lit), rtol=1e-13, atol=1e-300)
        np.testing.assert_array_equal(orc.forward(w, idx), np.array(S.slave_forward(model, data, wl, idx.tolist())))
        loss, acc = orc.loss_acc(w, idx=idx)
        batch = [data[i] for i in idx]
        assert np.isclose(loss, S.local_loss(model, wl, batch), rtol=1e-13)
        assert acc == S.local_accuracy(model, wl, batch)
        loss2, acc2 = orc.loss_acc(w, begin=n_train, n=n - n_train)
        assert np.isclose(loss2, S.local_loss(model, wl, data[n_train:]), rtol=1e-13)
        assert acc2 == S.local_accuracy(model, wl, data[n_train:])


@pytest.mark.parametrize("K,threads", [(1, 1), (2, 1), (3, 3), (4, 4)])
def test_sync_steps_trajectory(K, threads):
    rng = np.random.default_rng(100 + K)
    dim, n = 32, 80
    rp, col, val, lab = random_csr(rng, n, dim, max_nnz=6, dup_values=True)
    lam, lr, B, steps = 0.01, 0.5, 5, 12
    orc = Oracle(rp, col, val, lab, dim, lam)
    data = to_literal(rp, col, val, lab, dim)
    model = S.SparseSVM(lam, S.dim_sparsity(data))
    orc.set_dim_sparsity(orc.dim_sparsity(n))
    groups = S.split_vanilla(n, K)
    idx = np.stack([np.concatenate([rng.choice(g, size=B, replace=False) for g in groups]) for _ in range(steps)])
    wl = S.Sparse.zeros(dim)
    w = np.zeros(dim)
    losses_lit = []
    for s in range(steps):
        batches = [idx[s, k * B:(k + 1) * B].tolist() for k in range(K)]
        losses_lit.append(S.local_loss(model, wl, [data[i] for b in batches for i in b]))
        wl = S.master_sync_step(model, data, wl, batches, lr)
    w, losses = orc.sync_steps(w, idx.reshape(-1), [B] * K, lr, n_steps=steps, threads=threads)
    np.testing.assert_allclose(w, literal_to_w(wl, dim), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(losses, losses_lit, rtol=1e-12)
    assert losses[0] == 1.0  # KA1


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_all_cores_context_form_agrees_with_the_serial_one(threads):
    """bench.py's `all_cores` context figure splits ONE worker's batch over T threads; it must compute the same
    trajectory as the one-thread form (to rounding: the batch sum is associated differently), supports included."""
    rng = np.random.default_rng(55)
    dim, n = 300, 2000
    rp, col, val, lab = random_csr(rng, n, dim, max_nnz=12, dup_values=True)
    lam, lr, B, steps = 0.01, 0.5, 64, 40
    orc = Oracle(rp, col, val, lab, dim, lam)
    orc.set_dim_sparsity(orc.dim_sparsity(n))
    idx = np.stack([rng.choice(n, size=B, replace=False) for _ in range(steps)]).astype(np.int32).reshape(-1)
    w0 = rng.standard_normal(dim) * (rng.random(dim) < 0.5) * 0.1
    w1, l1 = orc.sync_steps(w0, idx, [B], lr, n_steps=steps)
    w2, l2 = orc.sync_steps_allcores(w0, idx, B, lr, steps, threads)
    np.testing.assert_allclose(l2, l1, rtol=1e-12)
    np.testing.assert_allclose(w2, w1, rtol=1e-10, atol=1e-14)
    if threads == 1:
        assert np.array_equal(w1, w2) and np.array_equal(l1, l2)


def test_async_delta_and_run():
    rng = np.random.default_rng(7)
    dim, n = 24, 50
    rp, col, val, lab = random_csr(rng, n, dim, max_nnz=6, dup_values=True)
    lam, lr = 0.02, 0.5
    orc = Oracle(rp, col, val, lab, dim, lam)
    data = to_literal(rp, col, val, lab, dim)
    model = S.SparseSVM(lam, S.dim_sparsity(data))
    orc.set_dim_sparsity(orc.dim_sparsity(n))
    for batch in (1, 4):
        idx = rng.integers(0, n, size=(30, batch)).astype(np.int32)
        wl = S.Sparse.zeros(dim)
        for u in range(len(idx)):
            delta = S.async_worker_delta(model, data, wl, idx[u].tolist(), lr)
            np.testing.assert_allclose(orc.async_delta(literal_to_w(wl, dim), idx[u], lr), literal_to_w(delta, dim),
                                       rtol=1e-13, atol=0)
            wl = wl - delta  # Slave.scala:101
        w = orc.async_run(np.zeros(dim), idx.reshape(-1), batch, lr)
        np.testing.assert_allclose(w, literal_to_w(wl, dim), rtol=1e-12, atol=1e-15)


def test_error_codes():
    rng = np.random.default_rng(3)
    rp, col, val, lab = random_csr(rng, 10, 8, max_nnz=3)
    orc = Oracle(rp, col, val, lab, 8, 0.1)
    with pytest.raises(OracleError):
        orc.gradient(np.zeros(8), np.zeros(0, dtype=np.int32))      # Q7: empty batch
    with pytest.raises(OracleError):
        orc.gradient(np.zeros(8), np.array([10], dtype=np.int32))   # out of range
