"""Host-side N > 1 logic on CPU: world_size-2 gloo processes drive the SPMD Master with a stand-in device
context (records what would be sent to the GPU).  Checks that ranks draw identical batches, take disjoint
slices, and that the sharded evaluation plumbing sums exactly."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeCtx:
    """Stands in for NativeCtx: no arithmetic, just bookkeeping of the calls a rank would make."""

    def __init__(self, dim):
        self.dim, self.calls, self.w = dim, [], np.zeros(dim)

    def set_weights(self, w):
        self.w = np.array(w, dtype=np.float64)

    def get_weights(self):
        return self.w.copy()

    def set_workers(self, counts, k_total):
        self.calls.append(("workers", list(map(int, counts)), int(k_total)))

    def sync_steps(self, samples, n_per_step, n_steps, lr, want_losses=True):
        self.calls.append(("steps", np.array(samples).copy(), n_per_step, n_steps))
        return np.zeros(n_steps)

    def eval_counts(self, lo, hi, w=None):
        return (hi - lo) * 1, (hi - lo) // 2, 4.0          # hinge 1 per row, half correct, ||w||^2 = 4

    def comm_init(self, uid):
        self.calls.append(("comm", bytes(uid)))

    def close(self):
        pass


class FakeSlave:
    def __init__(self, rank, world, n_train, n_test, dim):
        self.ctx, self.world, self.is_async = FakeCtx(dim), world, False
        self.n_train, self.n_test, self.dim = n_train, n_test, dim


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from distributed_sgd_b200.core import Group, master as master_mod
    from distributed_sgd_b200.ml import EarlyStopping, SparseSVM
    from distributed_sgd_b200.utils.dataset import Data

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    # the unique id normally comes from libdsgd.so (NCCL); on CPU it is any 128-byte token
    master_mod.NativeCtx.comm_unique_id = staticmethod(lambda: bytes(range(128)))
    n_train, n_test, dim = 101, 40, 16
    stub = lambda n: Data(np.arange(n + 1, dtype=np.int64), np.zeros(n, np.int32), np.ones(n, np.float32),
                          np.ones(n, np.int8), dim)
    slave = FakeSlave(rank, world, n_train, n_test, dim)
    m = master_mod.MasterSync(rank, stub(n_train), stub(n_test), SparseSVM(0.5), world, slave=slave, group=Group(), seed=0)
    state = m.fit(np.zeros(dim), max_epochs=2, batch_size=10, learning_rate=0.5,
                  stopping_criterion=EarlyStopping.no_improvement(patience=5, min_delta=0.01), virtual_workers=2)
    steps = [c for c in slave.ctx.calls if c[0] == "steps"]
    workers = [c for c in slave.ctx.calls if c[0] == "workers"]
    comm = [c for c in slave.ctx.calls if c[0] == "comm"]
    q.put({"rank": rank, "samples": [s[1].tolist() for s in steps], "shapes": [(s[2], s[3]) for s in steps],
           "workers": [(w[1], w[2]) for w in workers], "comm": comm[0][1] if comm else None,
           "loss": m.history["losses"], "acc": m.history["accs"], "updates": state.updates})
    dist.destroy_process_group()


def test_spmd_master_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=30)
    r0, r1 = res
    assert r0["comm"] == r1["comm"] == bytes(range(128))            # rank 0's id reached rank 1
    # K = 4 logical workers over 101 rows: vanilla groups of 26, 26, 26, 23 -> ranks own groups {0,1} and {2,3}
    assert r0["workers"][0] == ([10, 10], 4) and r1["workers"][0] == ([10, 10], 4)
    flat0 = np.concatenate([np.array(s) for s in r0["samples"]])
    flat1 = np.concatenate([np.array(s) for s in r1["samples"]])
    assert flat0.min() >= 0 and flat0.max() < 52 and flat1.min() >= 52 and flat1.max() < 101   # disjoint slices
    # epoch = ceil(26 / 10) = 3 steps; the last step is ragged (6 rows for full groups, 3 for the short one)
    assert r0["shapes"][:2] == [(20, 2), (12, 1)] and r1["shapes"][:2] == [(20, 2), (9, 1)]
    assert r0["workers"][1] == ([6, 6], 4) and r1["workers"][1] == ([6, 3], 4)
    # sharded evaluation: hinge 1 per row -> loss = lambda * ||w||^2 + 1; accuracy = (50 // 2 + 51 // 2) / 101
    assert r0["loss"] == r1["loss"] == [0.5 * 4.0 + 1.0] * 2
    assert r0["acc"] == r1["acc"] == [(50 // 2 + 51 // 2) / 101] * 2
    assert r0["updates"] == 2


def test_ranks_issue_the_same_call_sequence_when_the_last_group_is_short():
    """n_train = 41, 2 ranks, batch 7 -> groups of 21 and 20 rows; steps draw (7,7), (7,7), (7,6).
    Rank 0's own counts never change, rank 1's do: cutting the epoch into device calls by a rank's OWN counts made the ranks
    issue different call sequences (the fused multi-GPU kernel numbers its exchange by call and would spin into its
    watchdog).  The boundaries must come from the global shape."""
    sys.path.insert(0, ROOT)
    from distributed_sgd_b200.core import master as master_mod
    from distributed_sgd_b200.ml import EarlyStopping, SparseSVM
    from distributed_sgd_b200.utils.dataset import Data

    class OneRankOfTwo:           # a Group that claims rank r of 2 without a process group
        def __init__(self, r):
            self.rank, self.world, self.active = r, 2, False
        def barrier(self): pass
        def broadcast_bytes(self, b, src=0): return b
        def all_gather_bytes(self, b): return [b, b]
        def all_reduce_sum(self, v): return list(v)
        def all_reduce_max(self, v): return v

    n_train, n_test, dim = 41, 10, 8
    stub = lambda n: Data(np.arange(n + 1, dtype=np.int64), np.zeros(n, np.int32), np.ones(n, np.float32),
                          np.ones(n, np.int8), dim)
    shapes = []
    for r in range(2):
        slave = FakeSlave(r, 2, n_train, n_test, dim)
        m = master_mod.MasterSync(r, stub(n_train), stub(n_test), SparseSVM(0.5), 2, slave=slave, group=OneRankOfTwo(r),
                                  seed=0, attach=False)
        m.fit(np.zeros(dim), max_epochs=1, batch_size=7, learning_rate=0.5,
              stopping_criterion=EarlyStopping.no_improvement(patience=5, min_delta=0.01))
        shapes.append([(c[2], c[3]) for c in slave.ctx.calls if c[0] == "steps"])
    assert shapes[0] == [(7, 2), (7, 1)] and shapes[1] == [(7, 2), (6, 1)]      # same number of calls, same step counts


def test_empty_slice_fails_the_fit_on_every_rank():
    """Quirk Q7 (math/Vec.scala:129 via core/Master.scala:187): a worker whose slice is empty at some step makes Vec.sum
    throw; every rank must notice at that step, whichever rank owns the empty slice."""
    sys.path.insert(0, ROOT)
    import pytest
    from distributed_sgd_b200.core import master as master_mod
    from distributed_sgd_b200.ml import EarlyStopping, SparseSVM
    from distributed_sgd_b200.utils.dataset import Data
    n_train, dim = 29, 8                          # 2 workers: groups of 15 and 14; batch 7 -> third step draws (1, 0)
    stub = lambda n: Data(np.arange(n + 1, dtype=np.int64), np.zeros(n, np.int32), np.ones(n, np.float32),
                          np.ones(n, np.int8), dim)
    slave = FakeSlave(0, 1, n_train, 5, dim)
    m = master_mod.MasterSync(0, stub(n_train), stub(5), SparseSVM(0.5), 1, slave=slave, seed=0)
    with pytest.raises(ValueError, match="empty list"):
        m.fit(np.zeros(dim), max_epochs=1, batch_size=7, learning_rate=0.5,
              stopping_criterion=EarlyStopping.no_improvement(patience=5, min_delta=0.01), virtual_workers=2)
