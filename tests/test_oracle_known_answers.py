"""Pins the literal restatement (oracle/scala_semantics.py) against every known answer the reference
holds or implies for this path (VecTests + KA1..KA9)."""
import math

import pytest

from oracle import scala_semantics as S
from oracle.scala_semantics import Sparse, SparseSVM


# ---- KA8: the reference's own VecTests (src/test/scala/epfl/distributed/data/VecTests.scala) ----------

def test_vectests_sparse_add():  # VecTests.scala:25-30
    v1 = Sparse({0: 1, 1: 2, 2: 3}, 4)
    v2 = Sparse({1: 1, 2: 2, 3: 3}, 4)
    assert v1 + v2 == Sparse({0: 1, 1: 3, 2: 5, 3: 3}, 4)


def test_vectests_dense_cases_on_sparse_algebra():  # VecTests.scala:12-23 (same algebra, sparse carrier)
    v1 = Sparse({0: 1, 1: 2, 2: 3}, 3)
    v2 = Sparse({0: 1, 1: 2, 2: 3}, 3)
    assert v1 + v2 == Sparse({0: 2, 1: 4, 2: 6}, 3)
    assert v1.dot(v2) == 1 + 4 + 9
    assert v1 * 2 == Sparse({0: 2, 1: 4, 2: 6}, 3)
    assert 3 * v1 == Sparse({0: 3, 1: 6, 2: 9}, 3)
    assert v1.norm() == math.sqrt(1 + 4 + 9)


def test_vectests_division_by_zero_throws():  # VecTests.scala:32-35
    with pytest.raises(ValueError):
        Sparse({0: 1, 1: 2, 2: 3}, 4) / 0


def test_vectests_sparsity():  # VecTests.scala:37-41
    assert Sparse({0: 1, 1: 2}, 10).sparsity() == 0.8


# ---- known answers derivable from the source ----------------------------------------------------------

def _toy():
    data = [(Sparse({1: 1.0, 3: 0.5}, 8), 1), (Sparse({2: 2.0, 3: 0.25}, 8), -1), (Sparse({1: 0.5, 4: 1.0}, 8), -1)]
    d = S.dim_sparsity(data)
    return data, SparseSVM(1e-5, d)


def test_ka1_zero_weights_loss_is_one_accuracy_zero():  # SparseSVM.scala:14,16; Master.scala:102
    data, model = _toy()
    w0 = Sparse.zeros(8)
    assert all(model.forward(w0, x) == 0 for x, _ in data)
    assert S.local_loss(model, w0, data) == 1.0
    assert S.local_accuracy(model, w0, data) == 0.0


def test_ka2_zero_weights_gradient_is_sum_yx():  # SparseSVM.scala:28,31; Vec.scala:66
    data, model = _toy()
    g = S.slave_gradient(model, data, Sparse.zeros(8), [0, 1, 2])
    assert g == Sparse({1: 1.0 - 0.5, 3: 0.5 - 0.25, 2: -2.0, 4: -1.0}, 8)


def test_ka3_gate_is_inclusive_at_zero():  # SparseSVM.scala:28
    model = SparseSVM(0.0, Sparse.zeros(4))
    x = Sparse({1: 1.0, 2: 1.0}, 4)
    w = Sparse({1: 1.0, 2: -1.0}, 4)  # x.w == 0 exactly
    assert model.backward(w, x, 1) == x * 1
    assert model.backward(w, x, -1) == x * -1


def test_ka4_single_row_two_steps():
    data = [(Sparse({1: 1.0}, 4), 1)]
    model = SparseSVM(1e-5, S.dim_sparsity(data))
    w1 = S.master_sync_step(model, data, Sparse.zeros(4), [[0]], 0.5)
    assert w1 == Sparse({1: -0.5}, 4)
    assert model.forward(w1, data[0][0]) == 1.0 and model.loss_sample(w1, *data[0]) == 0.0
    g2 = S.slave_gradient(model, data, w1, [0])
    assert g2 == Sparse.zeros(4)  # activity = -0.5 < 0: empty support, no +c anywhere
    assert S.master_sync_step(model, data, w1, [[0]], 0.5) == w1


def test_ka5_exact_cancellation_drops_the_key_and_gets_no_c():  # Sparse.scala:108-118 (quirk Q9)
    data = [(Sparse({1: 0.5, 2: 1.0}, 6), 1), (Sparse({1: 0.5, 3: 1.0}, 6), -1)]
    d = Sparse({0: 0.5, 1: 0.5, 2: 0.5, 3: 0.5}, 6)
    model = SparseSVM(0.1, d)
    w = Sparse({2: 0.25, 3: -0.25, 5: 1.0}, 6)  # both activities > 0; w.d = 0.125 - 0.125 = 0 -> make it non-zero:
    w = Sparse({2: 0.25, 3: -0.125, 5: 1.0}, 6)
    c = 0.1 * 2.0 * w.dot(d)
    assert c != 0
    g = S.slave_gradient(model, data, w, [0, 1])
    assert 1 not in g.map                       # cancelled key absent: no regulariser there
    assert g.map[2] == 1.0 + c and g.map[3] == -1.0 + c


def test_ka6_no_improvement():  # EarlyStopping.scala:30-42
    crit = S.early_stopping_no_improvement(patience=5, min_delta=0.01)
    assert crit([]) is False
    assert crit([0.5, 0.6, 0.7]) is False                        # min is the newest
    assert crit([0.9, 0.9, 0.9, 0.9, 0.9, 0.5]) is True          # best is 5 back
    assert crit([0.9, 0.9, 0.9, 0.9, 0.5]) is False              # best is 4 back < patience
    # tolerance scan: an older value within |minDelta| of the running min takes over the arg-min
    assert crit([0.500, 0.9, 0.9, 0.9, 0.9, 0.505]) is True
    assert S.early_stopping_target(0.3)([0.2, 0.9]) is True and S.early_stopping_target(0.3)([0.4]) is False


def test_ka7_vanilla_split():  # SplitStrategy.scala:13-14
    assert [len(g) for g in S.split_vanilla(10, 4)] == [3, 3, 3, 1]
    assert [len(g) for g in S.split_vanilla(9, 4)] == [3, 3, 3]
    assert S.split_vanilla(6, 2) == [[0, 1, 2], [3, 4, 5]]


def test_ka9_sync_step_counts_supports_per_worker():  # Master.scala:194,197; Vec.scala:72 (H4)
    data = [(Sparse({1: 1.0, 2: 1.0}, 6), 1), (Sparse({2: 1.0, 3: 1.0}, 6), 1)]
    d = Sparse({1: 0.5, 2: 0.5, 3: 0.5}, 6)
    model = SparseSVM(0.1, d)
    w = Sparse({1: 0.5, 2: 0.5, 3: 0.5}, 6)
    c = 0.1 * 2.0 * w.dot(d)
    lr, K = 0.5, 2
    w1 = S.master_sync_step(model, data, w, [[0], [1]], lr)
    assert w1.map[1] == 0.5 - lr * ((1.0 + c) / K)
    assert w1.map[2] == 0.5 - lr * (((1.0 + c) + (1.0 + c)) / K)   # both supports contain key 2 -> 2c
    assert w1.map[3] == 0.5 - lr * ((1.0 + c) / K)


def test_q7_empty_batch_throws():  # Vec.scala:129
    data, model = _toy()
    with pytest.raises(ValueError):
        S.slave_gradient(model, data, Sparse.zeros(8), [])


def test_q3_dim_sparsity_keys_are_shifted():  # Main.scala:60,62
    data = [(Sparse({1: 1.0, 3: 1.0}, 8), 1), (Sparse({3: 1.0}, 8), -1)]
    d = S.dim_sparsity(data)
    assert d == Sparse({0: 1.0 / 2, 2: 1.0 / 3}, 8)


def test_async_delta_is_mean_then_regularize_then_lr():  # Slave.scala:92-99 (Q4)
    data = [(Sparse({1: 1.0}, 4), 1), (Sparse({2: 1.0}, 4), 1)]
    d = Sparse({1: 1.0}, 4)
    model = SparseSVM(0.25, d)
    w = Sparse({1: 2.0}, 4)
    c = 0.25 * 2.0 * 2.0
    delta = S.async_worker_delta(model, data, w, [0, 1], 0.5)
    assert delta == Sparse({1: 0.5 * (0.5 + c), 2: 0.5 * (0.5 + c)}, 4)
