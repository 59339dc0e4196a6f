"""Size-independent properties at BASELINE.json's full size (47 236 features, 700 000 rows, ~0.2 % non-zeros) and the
Main.scala calling sequence end to end.  (File name sorts last on purpose: these take the longest.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    from distributed_sgd_b200.native import NativeCtx
    from distributed_sgd_b200.utils import synthetic_rcv1
    data = synthetic_rcv1(n_rows=700_000, seed=0)
    n_train = 560_000                                                       # Main.scala:52
    ctx = NativeCtx(0, data.dim, 1e-5)
    ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
    ctx.compute_dim_sparsity(n_train)
    yield data, n_train, ctx
    ctx.close()


def test_zero_weights_known_answer_at_full_size(full):
    """KA1 on every row: w = 0 => every prediction 0, loss exactly 1, accuracy exactly 0 (SparseSVM.scala:14,16)."""
    data, n_train, ctx = full
    w0 = np.zeros(data.dim)
    assert ctx.eval(0, n_train, w0) == (1.0, 0.0)
    assert ctx.eval(n_train, data.n_rows, w0) == (1.0, 0.0)
    h, c, n2 = ctx.eval_counts(0, data.n_rows, w0)
    assert (h, c, n2) == (data.n_rows, 0, 0.0)


def test_gradient_is_additive_over_batches_at_zero_weights(full):
    """KA2 + linearity: at w = 0 nothing is gated and c = 0, so gradient(A u B) = gradient(A) + gradient(B) = sum y x."""
    data, n_train, ctx = full
    rng = np.random.default_rng(0)
    perm = rng.permutation(n_train)[:300_000].astype(np.int32)
    a, b = perm[:170_000], perm[170_000:]
    w0 = np.zeros(data.dim)
    ga, gb, gab = ctx.gradient(a, w0), ctx.gradient(b, w0), ctx.gradient(perm, w0)
    np.testing.assert_allclose(gab, ga + gb, rtol=1e-12, atol=1e-10)
    # checksum of the whole batch against plain numpy on the host: sum_j g_j = sum_i y_i * sum_j x_ij
    row_sums = np.add.reduceat(data.val.astype(np.float64), data.row_ptr[:-1])
    expect = float(np.sum(data.label[perm].astype(np.float64) * row_sums[perm]))
    assert float(gab.sum()) == pytest.approx(expect, rel=1e-9, abs=1e-6)
    # per-column check on a slice of the columns, also against numpy
    sel = np.zeros(data.n_rows, dtype=bool); sel[perm] = True
    rows_of = np.repeat(np.arange(data.n_rows), np.diff(data.row_ptr))
    mask = sel[rows_of]
    ref = np.bincount(data.col[mask], weights=data.val[mask].astype(np.float64) * data.label[rows_of[mask]], minlength=data.dim)
    np.testing.assert_allclose(gab, ref, rtol=1e-11, atol=1e-10)


def test_eval_counts_add_over_row_shards(full):
    data, n_train, ctx = full
    rng = np.random.default_rng(1)
    w = np.where(rng.random(data.dim) < 0.5, rng.standard_normal(data.dim) * 0.05, 0.0)
    cuts = [0, 1, 70_000, 70_001, 333_333, n_train]
    parts = [ctx.eval_counts(lo, hi, w) for lo, hi in zip(cuts, cuts[1:])]
    h, c, n2 = ctx.eval_counts(0, n_train, w)
    assert (sum(p[0] for p in parts), sum(p[1] for p in parts)) == (h, c)
    assert all(p[2] == n2 for p in parts)                                   # same fixed-order reduction every time
    # predictions through forward() on a strided sample agree with the counters
    idx = np.arange(0, n_train, 97, dtype=np.int32)
    preds = ctx.forward(idx, w)
    assert int(np.sum(preds == data.label[idx])) == sum(ctx.eval_counts(int(i), int(i) + 1, w)[1] for i in idx[:50]) + \
        int(np.sum(preds[50:] == data.label[idx[50:]]))


def test_sync_epoch_at_full_size_is_reproducible(full):
    """One reference epoch of the 1-worker fit loop at the bench configuration (2188 steps of batch 256): the run split
    into two calls gives the same trajectory to fp64 rounding, and the resident-weights and request-weights evaluation
    paths agree exactly."""
    data, n_train, ctx = full
    rng = np.random.default_rng(2)
    B, S = 256, 2188
    idx = np.stack([rng.choice(n_train, size=B, replace=False) for _ in range(S)]).astype(np.int32).reshape(-1)
    ctx.set_weights(np.zeros(data.dim))
    l1 = ctx.sync_steps(idx, B, S, 0.5)
    w1 = ctx.get_weights()
    ctx.set_weights(np.zeros(data.dim))
    la = ctx.sync_steps(idx[:B * 1000], B, 1000, 0.5)
    lb = ctx.sync_steps(idx[B * 1000:], B, S - 1000, 0.5)
    w2 = ctx.get_weights()
    assert l1[0] == 1.0
    np.testing.assert_allclose(np.concatenate([la, lb]), l1, rtol=1e-9)
    np.testing.assert_allclose(w2, w1, rtol=1e-8, atol=1e-12)
    loss, acc = ctx.eval(n_train, data.n_rows)
    loss_r, acc_r = ctx.eval(n_train, data.n_rows, w2)                      # resident vs request weights:
    assert acc == acc_r and loss == pytest.approx(loss_r, rel=1e-12)        # ||w||^2 is reduced in a different (fixed) order
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0 and np.all(np.isfinite(w2)) and np.count_nonzero(w2) > 10_000


def test_configs0_small_slice_application_conf_defaults_match_oracle():
    """BASELINE.json configs[0]: sync mode, 2 workers, application.conf defaults (batch 100, lr 0.5, lambda 1e-5, 10 epochs,
    patience 5, conv-delta 0.01) on a 23 149-row slice (the size of RCV1's train file, utils/Dataset.scala:47-50), through
    Main.scenario -- every per-epoch list, the stopping epoch and the final weights against the oracle driven with the same
    batch draws (core/Master.scala:140-213; Main.scala:70-120)."""
    from distributed_sgd_b200.main import scenario
    from distributed_sgd_b200.utils import load_config, synthetic_rcv1
    from distributed_sgd_b200.ml import EarlyStopping
    from oracle.oracle import Oracle
    data = synthetic_rcv1(n_rows=23149, seed=5)
    cfg = load_config(env={"DSGD_NODE_COUNT": "2"})
    assert (cfg.batch_size, cfg.learning_rate, cfg.lam, cfg.max_epochs, cfg.patience, cfg.conv_delta) == (100, 0.5, 1e-5, 10, 5, 0.01)
    drawn, final = [], {}

    def inspect(what, obj):
        if what == "master":
            orig = obj.draw_epoch
            obj.draw_epoch = lambda groups, bs, *a: drawn.append(orig(groups, bs, *a)) or drawn[-1]
        else:
            final["master"], final["state"] = obj

    rep = scenario(cfg, data, rank=0, world=1, device=0, log=lambda s: None, inspect=inspect)
    n_train = int(data.n_rows * 0.8)
    orc = Oracle(data.row_ptr, data.col, data.val, data.label, data.dim, cfg.lam)
    orc.set_dim_sparsity(orc.dim_sparsity(n_train))
    assert rep["initial_loss"] == 1.0 and rep["initial_accuracy"] == 0.0    # w0 = 0 (Main.scala:74-78)
    crit = EarlyStopping.no_improvement(patience=cfg.patience, min_delta=cfg.conv_delta, min_steps=None)
    w = np.zeros(data.dim)
    losses, accs, tlosses, taccs = [], [], [], []
    epochs = 0
    while not (epochs >= cfg.max_epochs or crit(tlosses[::-1])):            # Master.scala:154,166 (newest first)
        for step in drawn[epochs]:
            assert [len(b) for b in step][0] in (cfg.batch_size, 9260 % cfg.batch_size)
            w, _ = orc.sync_steps(w, np.concatenate(step), [len(b) for b in step], cfg.learning_rate, n_steps=1)
        l, a = orc.loss_acc(w, begin=0, n=n_train); losses.append(l); accs.append(a)
        l, a = orc.loss_acc(w, begin=n_train, n=data.n_rows - n_train); tlosses.append(l); taccs.append(a)
        epochs += 1
    h = rep["history"]
    assert rep["updates"] == epochs and len(h["losses"]) == epochs
    np.testing.assert_allclose(h["losses"], losses, rtol=1e-11)
    np.testing.assert_allclose(h["test_losses"], tlosses, rtol=1e-11)
    assert h["accs"] == accs and h["test_accs"] == taccs
    np.testing.assert_allclose(final["state"].grad, w, rtol=1e-10, atol=1e-15)
    assert rep["final_test_loss"] == pytest.approx(tlosses[-1], rel=1e-11) and rep["final_test_accuracy"] == taccs[-1]


def test_main_scenario_async_replays_through_the_literal_master_loop():
    """Main.scenario in async mode (Main.scala:82-96): the run itself is a race (Hogwild), so the check is a replay: the
    polls the master made (update counter, raw test loss / accuracy of the snapshot it took) go through the literal
    restatement of core/MasterAsync.scala:96-177 and must give the same leaky lists, the same best snapshot, the same end."""
    from distributed_sgd_b200.main import scenario
    from distributed_sgd_b200.utils import load_config, synthetic_rcv1
    from oracle import scala_semantics as S
    data = synthetic_rcv1(n_rows=6000, seed=5)
    cfg = load_config(env={"DSGD_ASYNC": "true", "DSGD_BATCH_SIZE": "1", "DSGD_MAX_EPOCHS": "5", "DSGD_CHECK_EVERY": "2000",
                           "DSGD_LEARNING_RATE": "0.1"})
    final = {}

    def inspect(what, obj):
        final[what] = obj
        if what == "done":      # the device context is still alive here: re-evaluate the returned weights
            master, state = obj
            final["loss_of_returned_weights"] = master.local_loss_accuracy(state.grad, test_data=True)[0]

    rep = scenario(cfg, data, rank=0, world=1, device=0, log=lambda s: None, async_concurrency=8, inspect=inspect)
    master, state = final["done"]
    h = master.history
    assert rep["initial_loss"] == 1.0 and len(h["test_losses"]) >= 1
    raw = iter(zip(h["raw_test_losses"], h["raw_test_accs"]))
    polls = []
    for i, (u, computed) in enumerate(h["polls"]):
        l, a = next(raw) if computed else (None, None)
        polls.append((u, l, a, i))
    ref = S.MasterAsyncLossChecker(int(data.n_rows * 0.8), cfg.max_epochs,
                                   S.early_stopping_no_improvement(cfg.patience, cfg.conv_delta), cfg.check_every,
                                   cfg.leaky_loss).replay(polls)
    assert ref["computed_at"] == h["checks_at"] and ref["ended_by"] == h["ended_by"]
    assert ref["test_losses"] == h["test_losses"] and ref["test_accs"] == h["test_accs"]
    assert state.loss == ref["best_loss"] == min(h["test_losses"])
    # the returned weights are the snapshot of the best check: re-evaluating them gives that check's RAW test loss
    assert final["loss_of_returned_weights"] == pytest.approx(h["raw_test_losses"][h["best_check"]], rel=1e-12)
    assert rep["final_test_accuracy"] > 0.5 and rep["final_weights_nonzero"] > 0


# ---- golden fixtures (tests/golden/*.json, from the literal restatement of the Scala arithmetic) through the CUDA path ----
from test_golden import FIXTURES, flat_draws, load  # noqa: E402
import os  # noqa: E402


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cuda_path_reproduces_golden(path):
    from distributed_sgd_b200.native import NativeCtx
    f = load(path)
    ctx = NativeCtx(0, f["dim"], f["lambda"])
    ctx.load_csr(f["row_ptr"], f["col"], f["val"], f["label"])
    d = ctx.compute_dim_sparsity(f["n_train"])
    np.testing.assert_array_equal(d, f["dim_sparsity_weight_space"])
    ctx.set_weights(np.zeros(f["dim"]))
    ctx.set_workers([f["B"]] * f["K"], f["K"])                                               # K logical workers on one GPU
    losses = ctx.sync_steps(flat_draws(f), f["B"] * f["K"], len(f["draws"]), f["lr"])
    np.testing.assert_allclose(losses, f["step_losses"], rtol=1e-12)
    w = ctx.get_weights()
    np.testing.assert_allclose(w, f["final_weights"], rtol=1e-11, atol=1e-15)
    g = ctx.gradient(f["probe"], np.array(f["final_weights"]))
    assert (g == 0).tolist() == (np.array(f["probe_gradient"]) == 0).tolist()
    np.testing.assert_allclose(g, f["probe_gradient"], rtol=1e-12, atol=0)
    np.testing.assert_array_equal(ctx.forward(f["probe"], np.array(f["final_weights"])), f["probe_predictions"])
    n = len(f["label"])
    loss, acc = ctx.eval(f["n_train"], n, np.array(f["final_weights"]))
    assert acc == f["test_accuracy"] and loss == pytest.approx(f["test_loss"], rel=1e-12)
    ctx.close()
    actx = NativeCtx(0, f["dim"], f["lambda"], is_async=True)
    actx.load_csr(f["row_ptr"], f["col"], f["val"], f["label"])
    actx.set_dim_sparsity(np.array(f["dim_sparsity_weight_space"]))
    actx.async_replay(np.zeros(f["dim"]), np.array(f["async_samples"], np.int32), 1, f["lr"])
    np.testing.assert_allclose(actx.get_weights(), f["async_final_weights"], rtol=1e-9, atol=1e-13)
    actx.close()


@pytest.mark.parametrize("batch", [256, 1500])
def test_persistent_loop_with_rows_that_overflow_the_tma_stage(batch):
    """Very long rows (1800-2000 non-zeros each): with two rows per CTA the second one no longer fits the 20 KB
    shared-memory stage and is read from global memory chunk by chunk (batch 256); with ~10 rows per CTA the chunk
    list overflows too and whole rows take the warp-per-row slow path (batch 1500).  Same trajectory as the oracle."""
    from distributed_sgd_b200.native import NativeCtx
    from oracle.oracle import Oracle
    rng = np.random.default_rng(batch)
    dim, n = 47236, 1600
    lens = rng.integers(1800, 2001, size=n)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    col = np.concatenate([np.sort(rng.choice(dim, size=int(l), replace=False)) for l in lens]).astype(np.int32)
    val = (np.abs(rng.standard_normal(len(col))) * 0.02 + 1e-3).astype(np.float32)
    lab = rng.choice(np.array([-1, 1], dtype=np.int8), size=n)
    lam, lr, steps = 1e-3, 0.05, 6
    ctx = NativeCtx(0, dim, lam)
    ctx.load_csr(rp, col, val, lab)
    orc = Oracle(rp, col, val, lab, dim, lam)
    d = orc.dim_sparsity(n)
    orc.set_dim_sparsity(d)
    ctx.set_dim_sparsity(d)
    idx = np.stack([rng.choice(n, size=batch, replace=False) for _ in range(steps)]).astype(np.int32).reshape(-1)
    w_ref, losses_ref = orc.sync_steps(np.zeros(dim), idx, [batch], lr, n_steps=steps)
    ctx.set_weights(np.zeros(dim))
    losses = ctx.sync_steps(idx, batch, steps, lr)
    np.testing.assert_allclose(losses, losses_ref, rtol=1e-12)
    np.testing.assert_allclose(ctx.get_weights(), w_ref, rtol=1e-10, atol=1e-14)
    ctx.close()
