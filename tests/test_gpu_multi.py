"""Two-GPU parity of the sync step: one process per GPU, gradients summed on the devices (NCCL over NVLink
inside libdsgd.so), compared with the oracle's K = 2 master step.  Skipped on a single-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, mode):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from distributed_sgd_b200.core import Group
    from distributed_sgd_b200.native import NativeCtx
    from distributed_sgd_b200.utils import synthetic_rcv1
    from oracle.oracle import Oracle

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    group = Group()
    data = synthetic_rcv1(n_rows=6000, seed=3)
    n_train, lam, lr, batch, steps = 4800, 0.01, 0.5, 48, 20
    V = 2 if mode == "nccl" else 1      # the fused peer-memory kernel runs one worker per GPU
    ctx = NativeCtx(rank, data.dim, lam, rank=rank, world=world)
    ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
    d = ctx.compute_dim_sparsity(n_train)
    uid = NativeCtx.comm_unique_id() if rank == 0 else b""
    ctx.comm_init(group.broadcast_bytes(uid, 0))
    if mode == "p2p":
        ctx.setup_peer_exchange(group)
    rng = np.random.default_rng(5)                       # same stream on every rank, like Random.setSeed(0)
    K = world * V                                         # V logical workers per GPU
    per = n_train // K
    idx = np.stack([np.concatenate([k * per + rng.choice(per, size=batch, replace=False) for k in range(K)])
                    for _ in range(steps)]).astype(np.int32)          # [steps, K * batch]
    mine = idx.reshape(steps, K, batch)[:, rank * V:(rank + 1) * V, :].reshape(steps, V * batch)
    ctx.set_weights(np.zeros(data.dim))
    ctx.set_workers([batch] * V, K)
    losses = ctx.sync_steps(mine.reshape(-1), V * batch, steps, lr)
    if mode == "p2p":   # a second call continues from the first one's state (global step counter, buffers)
        half = steps // 2
        ctx.set_weights(np.zeros(data.dim))
        l1 = ctx.sync_steps(mine[:half].reshape(-1), V * batch, half, lr)
        l2 = ctx.sync_steps(mine[half:].reshape(-1), V * batch, steps - half, lr)
        assert np.array_equal(np.concatenate([l1, l2]), losses), "split run differs"
        launches = ctx.launch_count()
    w = ctx.get_weights()
    orc = Oracle(data.row_ptr, data.col, data.val, data.label, data.dim, lam)
    orc.set_dim_sparsity(d)
    w_ref, losses_ref = orc.sync_steps(np.zeros(data.dim), idx.reshape(-1), [batch] * K, lr, n_steps=steps)
    ok = bool(np.allclose(losses, losses_ref, rtol=1e-12, atol=0) and np.allclose(w, w_ref, rtol=1e-11, atol=1e-15))
    # replicas must be bit-identical across GPUs
    blobs = group.all_gather_bytes(w.tobytes())
    same = all(b == blobs[0] for b in blobs)
    if mode == "p2p":
        ok = ok and bool(np.array_equal(ctx.get_weights(), w))
    q.put((rank, ok, same, float(np.abs(w - w_ref).max())))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_two_gpu_sync_matches_oracle(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, same, err in res:
        assert ok, f"rank {rank}: trajectory differs from the oracle (max abs err {err})"
        assert same, "weight replicas differ across GPUs"
