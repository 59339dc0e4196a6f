"""Async (Hogwild) mode through the C ABI: deterministic replay against the oracle, the free-running device
loop, peer / master replicas, and the reference's mode errors (core/Slave.scala:159-195)."""
import time

import numpy as np
import pytest

from helpers import make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def synth():
    from distributed_sgd_b200.utils import synthetic_rcv1
    return synthetic_rcv1(n_rows=5000, seed=9)


@pytest.mark.parametrize("batch,n_updates", [(1, 300), (4, 120), (32, 40)])
def test_async_replay_matches_oracle(synth, batch, n_updates):
    """One lane over a recorded sampling sequence == the K = 1 case of Slave.asyncTask.  Tolerance: the device
    keeps S = w.d incrementally (fp64), the oracle recomputes it every iteration: rtol 1e-9 on the weights."""
    rng = np.random.default_rng(batch)
    ctx, orc = make_pair(synth, lam=1e-3, n_train=4000, is_async=True)
    idx = rng.integers(0, 4000, size=(n_updates, batch)).astype(np.int32)
    if batch > 1:  # without replacement inside a batch, like `shuffle take batchSize`
        idx = np.stack([rng.choice(4000, size=batch, replace=False) for _ in range(n_updates)]).astype(np.int32)
    w0 = np.zeros(synth.dim)
    ctx.async_replay(w0, idx.reshape(-1), batch, 0.5)
    w = ctx.get_weights()
    w_ref = orc.async_run(w0, idx.reshape(-1), batch, 0.5)
    assert (w == 0).tolist() == (w_ref == 0).tolist()
    np.testing.assert_allclose(w, w_ref, rtol=1e-9, atol=1e-13)
    assert ctx.async_updates() == n_updates
    ctx.close()


def test_async_free_running_with_master_replica(synth):
    ctx, orc = make_pair(synth, lam=1e-5, n_train=4000, is_async=True)
    w0 = np.zeros(synth.dim)
    ctx.async_host_master(w0)
    assigned = np.arange(4000, dtype=np.int32)
    loss0, _ = ctx.eval(4000, 5000, w0)
    ctx.start_async(w0, assigned, batch=1, lr=0.1, concurrency=1, max_updates=3000, seed=7)
    with pytest.raises(Exception):
        ctx.start_async(w0, assigned, batch=1, lr=0.1)              # "already running"
    t0 = time.time()
    while ctx.async_running() and time.time() - t0 < 60:
        time.sleep(0.01)
    assert not ctx.async_running()
    ctx.stop_async()
    assert ctx.async_updates() == 3000                              # the master counted every update
    w, wm = ctx.get_weights(), ctx.async_master_weights()
    np.testing.assert_array_equal(w, wm)                            # one lane: identical delta stream to both replicas
    loss1, acc1 = ctx.eval(4000, 5000, w)
    assert loss1 < loss0 and acc1 > 0.5
    # the loop body is the oracle's: replaying nothing else, c from S must track w.d
    ctx.close()


def test_async_hogwild_lanes_and_update_grad(synth):
    ctx, orc = make_pair(synth, lam=1e-5, n_train=4000, is_async=True)
    w0 = np.zeros(synth.dim)
    ctx.async_host_master(w0)
    ctx.start_async(w0, np.arange(4000, dtype=np.int32), batch=4, lr=0.05, concurrency=16, max_updates=4000, seed=3)
    # SlaveImpl.updateGrad while the loop runs: weights -= delta
    ctx.update_grad([5, 17], [0.25, -0.5])
    t0 = time.time()
    while ctx.async_running() and time.time() - t0 < 60:
        time.sleep(0.01)
    ctx.stop_async()
    assert ctx.async_updates() == 4000
    w, wm = ctx.get_weights(), ctx.async_master_weights()
    # the peer-pushed delta reached only this replica; everything else reached both (order differs: fp64 rounding)
    diff = w - wm
    assert diff[5] == pytest.approx(-0.25, abs=1e-12) and diff[17] == pytest.approx(0.5, abs=1e-12)
    diff[[5, 17]] = 0
    assert np.abs(diff).max() < 1e-12
    loss, acc = ctx.eval(4000, 5000, wm)
    assert acc > 0.5
    ctx.close()


def test_async_stop_interrupts_unbounded_loop(synth):
    ctx, orc = make_pair(synth, lam=1e-5, n_train=4000, is_async=True)
    ctx.start_async(np.zeros(synth.dim), np.arange(4000, dtype=np.int32), batch=1, lr=0.1, concurrency=4, max_updates=0, seed=1)
    time.sleep(0.05)
    assert ctx.async_running()
    n1 = ctx.async_updates()
    time.sleep(0.05)
    assert ctx.async_updates() > n1 > 0
    ctx.stop_async()
    assert not ctx.async_running()
    ctx.stop_async()                                                 # idempotent, like the reference
    ctx.close()


def test_async_mode_errors(synth):
    from distributed_sgd_b200 import native
    ctx, orc = make_pair(synth, lam=1e-5, n_train=4000, is_async=False)
    with pytest.raises(native.DsgdState):
        ctx.start_async(np.zeros(synth.dim), [0, 1], 1, 0.1)        # "slave is in synchronous mode"
    with pytest.raises(native.DsgdState):
        ctx.stop_async()
    ctx.close()
    actx, _ = make_pair(synth, lam=1e-5, n_train=4000, is_async=True)
    with pytest.raises(native.DsgdEmpty):
        actx.start_async(np.zeros(synth.dim), np.zeros(0, np.int32), 1, 0.1)
    with pytest.raises(native.DsgdRange):
        actx.start_async(np.zeros(synth.dim), [5000], 1, 0.1)
    actx.close()


def test_async_two_gpus_peer_writes(synth):
    """Two workers on two GPUs of one process: each pushes its deltas into the other's replica and the master's
    over NVLink (red.add on peer memory).  All three replicas end up equal up to fp64 summation order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    a, _ = make_pair(synth, lam=1e-5, n_train=4000, device=0, rank=0, world=2, is_async=True)
    b, _ = make_pair(synth, lam=1e-5, n_train=4000, device=1, rank=1, world=2, is_async=True)
    w0 = np.zeros(synth.dim)
    a.async_host_master(w0)
    from distributed_sgd_b200.native import REPLICA_MASTER, REPLICA_SELF
    a.peer_attach(1, b, REPLICA_SELF)
    b.peer_attach(0, a, REPLICA_SELF)
    b.peer_attach(2, a, REPLICA_MASTER)
    # initialise every replica first, then start the loops with w0 = None: a delta pushed by the faster worker
    # before the slower one starts must not be overwritten
    a.set_weights(w0); b.set_weights(w0)
    a.start_async(None, np.arange(0, 2000, dtype=np.int32), batch=1, lr=0.1, concurrency=8, max_updates=2000, seed=1)
    b.start_async(None, np.arange(2000, 4000, dtype=np.int32), batch=1, lr=0.1, concurrency=8, max_updates=2000, seed=2)
    t0 = time.time()
    while (a.async_running() or b.async_running()) and time.time() - t0 < 60:
        time.sleep(0.01)
    a.stop_async(); b.stop_async()
    assert a.async_updates() == 4000 == b.async_updates()
    wa, wb, wm = a.get_weights(), b.get_weights(), a.async_master_weights()
    assert np.abs(wa - wm).max() < 1e-12 and np.abs(wb - wm).max() < 1e-12
    assert a.eval(4000, 5000, wm)[1] > 0.5
    a.close(); b.close()


def test_master_async_fit_single_gpu(synth):
    """MasterAsync.fit end to end on one GPU: device loop + polling master logic + leaky loss + stop rule."""
    from distributed_sgd_b200 import MasterAsync, Slave, SparseSVM
    from distributed_sgd_b200.ml import EarlyStopping
    train, test = synth.split_at(4000)
    model = SparseSVM(1e-5)
    slave = Slave(0, 0, train, model, is_async=True, world=1, device=0, test_data=test)
    master = MasterAsync(0, train, test, model, 1, slave=slave)
    checks = []
    state = master.fit(np.zeros(synth.dim), max_epoch=40, batch_size=1, learning_rate=0.1,
                       stopping_criterion=EarlyStopping.no_improvement(patience=5, min_delta=0.01),
                       check_every=500, leak_loss_coef=0.9, concurrency=2, poll_seconds=0.0005,
                       on_check=lambda u, m: checks.append((u, m["test_loss"])))
    assert state.loss is not None and state.end is not None and state.updates == 1
    assert len(checks) >= 2 and all(b[0] - a[0] >= 500 for a, b in zip(checks, checks[1:]))
    assert state.loss == min(l for _, l in checks)                          # best smoothed loss is returned
    assert master.local_loss_accuracy(state.grad, test_data=True)[1] > 0.55
    assert not slave.ctx.async_running()
    slave.stop()


@pytest.mark.parametrize("batch", [1, 8])
def test_outbox_accumulates_exactly_what_the_worker_applied(synth, batch):
    """dsgd_async_outbox_*: the accumulator a host relay forwards to colleagues that are not GPU peers
    (core/Slave.scala:104-105) receives every -delta the worker applies to its own replica: after a recorded run,
    outbox == w_final - w0 up to the rounding of two sums that start from different values."""
    rng = np.random.default_rng(40 + batch)
    ctx, orc = make_pair(synth, lam=1e-3, n_train=4000, is_async=True)
    n_updates = 200
    idx = np.stack([rng.choice(4000, size=batch, replace=False) for _ in range(n_updates)]).astype(np.int32)
    w0 = rng.standard_normal(synth.dim) * (rng.random(synth.dim) < 0.2) * 0.05
    with pytest.raises(Exception):
        ctx.async_outbox_read()                                     # not enabled yet
    ctx.async_outbox_enable()
    assert not ctx.async_outbox_read().any()
    ctx.async_replay(w0, idx.reshape(-1), batch, 0.5)
    w, out = ctx.get_weights(), ctx.async_outbox_read()
    w_ref = orc.async_run(w0, idx.reshape(-1), batch, 0.5)
    np.testing.assert_allclose(w, w_ref, rtol=1e-9, atol=1e-13)     # the outbox does not disturb the run
    assert np.any(out != 0)
    np.testing.assert_allclose(out, w - w0, rtol=1e-12, atol=1e-15)
    ctx.close()


def test_async_deltas_relayed_to_a_grpc_colleague(synth):
    """The wire service with a colleague that is not a GPU peer: the colleague's replica (a plain numpy vector here) follows
    the GPU worker's through the relayed UpdateGrad messages."""
    from distributed_sgd_b200.core import wire

    class Colleague:
        dim = synth.dim

        def __init__(self, w0):
            self.w = w0.copy()
            self.n = 0

        def update_grad(self, idx, val):
            self.w[idx] -= val                                      # core/Slave.scala:180
            self.n += 1

    w0 = np.zeros(synth.dim)
    col = Colleague(w0)
    cserver, cport = wire.serve_slave(wire.SlaveServicer(col, n_train=4000, is_async=True), 0)
    ctx, _ = make_pair(synth, lam=1e-5, n_train=4000, is_async=True)
    srv = wire.SlaveServicer(ctx, n_train=4000, is_async=True, relay_period=0.02)
    M = srv.M
    try:
        srv.RegisterSlave(M.Node(host="127.0.0.1", port=cport))
        srv.StartAsync(M.StartAsyncRequest(weights=M.Sparse(size=synth.dim), samples=list(range(4000)), batchSize=1,
                                           learningRate=0.1))
        t0 = time.time()
        while ctx.async_updates() < 3000 and time.time() - t0 < 30:
            time.sleep(0.01)
        relay = srv.relay
        srv.StopAsync(M.Empty())
        assert ctx.async_updates() >= 3000 and relay.sent >= 1 and not relay.errors
        w = ctx.get_weights()
        assert np.any(w != 0)
        np.testing.assert_allclose(col.w, w, rtol=1e-10, atol=1e-14)   # same deltas, summed per period instead of one by one
    finally:
        cserver.stop(0)
    ctx.close()
