"""Dev tool: per-phase clock64 timeline of the persistent sync kernel (CTA 0), averaged over steps 50..250."""
import ctypes as C, os, sys
os.environ["DSGD_PERSIST_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributed_sgd_b200.native import NativeCtx, lib
from distributed_sgd_b200.utils import synthetic_rcv1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = 300
data = synthetic_rcv1(n_rows=200000, seed=0)
ctx = NativeCtx(0, data.dim, 1e-5)
ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
ctx.compute_dim_sparsity(160000)
rng = np.random.default_rng(0)
idx = np.stack([rng.choice(160000, size=B, replace=False) for _ in range(S)]).astype(np.int32).reshape(-1)
ctx.stage_samples(idx)
ctx.set_weights(np.zeros(data.dim))
for _ in range(2):
    ctx.sync_steps_staged(0, B, S, 0.5, want_losses=True)
ctx.synchronize()
tl_all = np.zeros(256 * 16 + 4 * 160 * 2, dtype=np.int64)
l = lib(); l.dsgd_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
assert l.dsgd_debug_timeline(ctx._h, tl_all.ctypes.data_as(C.c_void_p)) == 0
tl = tl_all[:4096].reshape(256, 16)
per_cta = tl_all[4096:].reshape(4, 160, 2)
names = {0: "interval start (consumer warp 0)", 1: "stage full (TMA landed)", 2: "pass 1 done (partial dots)",
         3: "pass 2 done (scatter issued)", 6: "CTA synced, arriving at grid barrier", 7: "grid barrier passed",
         8: "interval start (update warp 0)", 9: "c summed + handed over", 10: "update slice + partials published"}
t = tl[50:250]
base = t[:, 0:1]
print("batch", B, "CTAs", os.environ.get("DSGD_PERSIST_CTAS", "default"))
print("step period (cycles):", float(np.mean(np.diff(tl[50:250, 0]))))
for k in sorted(names):
    v = t[:, k] - base[:, 0]
    v = v[t[:, k] > 0]
    if len(v):
        print(f"  {names[k]:45s} +{np.mean(v):8.0f} cycles (min {v.min()}, max {v.max()})")

G = int(os.environ.get("DSGD_PERSIST_CTAS", "148"))
for k in range(4):
    a = per_cta[k, :G, 0]; b = per_cta[k, :G, 1]
    a0 = a.min()
    print(f"step {100+k} (ns, globaltimer): arrivals spread {a.max()-a0} (p50 {int(np.median(a-a0))}, p90 {int(np.percentile(a-a0,90))}); "
          f"last arrival -> first pass {b.min()-a.max()} ; last arrival -> last pass {b.max()-a.max()}")
