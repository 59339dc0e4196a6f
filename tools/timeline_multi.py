"""Dev tool: clock64 timeline of the fused multi-GPU sync kernel (CTA 0 of rank 0), 2+ processes over gloo."""
import ctypes as C, os, socket, sys
os.environ["DSGD_PERSIST_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def worker(rank, world, port, B):
    import numpy as np, torch, torch.distributed as dist
    from distributed_sgd_b200.core import Group
    from distributed_sgd_b200.native import NativeCtx, lib
    from distributed_sgd_b200.utils import synthetic_rcv1
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    group = Group()
    data = synthetic_rcv1(n_rows=200000, seed=0)
    ctx = NativeCtx(rank, data.dim, 1e-5, rank=rank, world=world)
    ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
    ctx.compute_dim_sparsity(160000)
    ctx.setup_peer_exchange(group)
    S = 300
    per = 160000 // world
    rng = np.random.default_rng(rank)
    idx = np.stack([rank * per + rng.choice(per, size=B, replace=False) for _ in range(S)]).astype(np.int32).reshape(-1)
    ctx.stage_samples(idx)
    ctx.set_weights(np.zeros(data.dim))
    for _ in range(3):
        group.barrier()
        ctx.timer_start()
        ctx.sync_steps_staged(0, B, S, 0.5, want_losses=True)
        ms = ctx.timer_stop()
    if rank == 0:
        tl = np.zeros(256 * 16 + 4 * 160 * 2, dtype=np.int64)
        l = lib(); l.dsgd_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
        assert l.dsgd_debug_timeline(ctx._h, tl.ctypes.data_as(C.c_void_p)) == 0
        tl = tl[:4096].reshape(256, 16)
        mode = int(os.environ.get("DSGD_P2P_MODE", "3"))
        if mode == 3:   # stamps of CTA 0 / warp 0 in the LL-word kernel
            names = {0: "interval start", 1: "push of g_{T-1} issued", 2: "c received", 3: "column update done (W_T word published)",
                     4: "stage full (rows landed)", 5: "pass 1 done (dots; waited on W_T words)", 8: "pass 2 done (scatter issued)",
                     6: "CTA synced, arriving at grid barrier", 7: "grid barrier passed"}
        else:
            names = {0: "step start", 1: "stage full", 2: "rows done (pass 2)", 3: "grid barrier 1 passed", 4: "push + flags issued",
                     5: "peer flags seen (+fence)", 8: "reduce + update slice done", 9: "partials published", 10: "grid barrier 2 passed"}
        t = tl[50:250]
        print(f"world {world} batch {B}: {ms*1e3/S:.2f} us/step; step period {np.mean(np.diff(tl[50:250,0])):.0f} cycles")
        for k in sorted(names):
            v = (t[:, k] - t[:, 0])[t[:, k] > 0]
            if len(v): print(f"  {names[k]:32s} +{np.mean(v):8.0f} cycles (min {v.min()}, max {v.max()})")
    group.barrier()
    ctx.close()
    dist.destroy_process_group()

if __name__ == "__main__":
    import torch.multiprocessing as mp
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.start_processes(worker, args=(world, port, B), nprocs=world, start_method="spawn")
