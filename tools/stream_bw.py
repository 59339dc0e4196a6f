"""Dev tool: achieved algorithmic GB/s of the streaming passes (eval over the train rows; large-batch gradient)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from distributed_sgd_b200.native import NativeCtx
from distributed_sgd_b200.utils import synthetic_rcv1
data = synthetic_rcv1(n_rows=700000, seed=0)
n_train = 560000
ctx = NativeCtx(0, data.dim, 1e-5)
ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
ctx.compute_dim_sparsity(n_train)
rng = np.random.default_rng(0)
ctx.set_weights(rng.standard_normal(data.dim) * 0.05)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
def best(fn, reps=5):
    fn(); ctx.synchronize()
    t = []
    for _ in range(reps):
        ctx.profile_begin(1); fn(); ms, n = ctx.profile_end(); t.append(ms)
    return min(t)
by = data.algorithmic_bytes(np.arange(n_train))
ms = best(lambda: ctx.eval(0, n_train))
print(f"eval {n_train} rows: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic = {by/ms/1e6/peak:.3f} of {peak} GB/s")
for n in (65536, 262144):
    idx = rng.choice(n_train, size=n, replace=False).astype(np.int32)
    by = data.algorithmic_bytes(idx)
    ms = best(lambda: ctx.gradient(idx))
    print(f"gradient batch {n}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic = {by/ms/1e6/peak:.3f} of peak")
    ms = best(lambda: ctx.forward(idx))
    print(f"forward  batch {n}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s algorithmic = {by/ms/1e6/peak:.3f} of peak")
