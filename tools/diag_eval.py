import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time
from distributed_sgd_b200.native import NativeCtx
from distributed_sgd_b200.utils import synthetic_rcv1
data = synthetic_rcv1(n_rows=100000, seed=0)
ctx = NativeCtx(0, data.dim, 1e-5)
ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
ctx.compute_dim_sparsity(80000)
rng = np.random.default_rng(0)
ctx.set_weights(rng.standard_normal(data.dim) * 0.05)
def t(fn):
    fn(); ctx.synchronize(); ctx.profile_begin(1); fn(); ms, n = ctx.profile_end(); return ms, n
for n in (2048, 4096, 65536):
    print("eval", n, t(lambda: ctx.eval(0, n)), flush=True)
    print("fwd-contig", n, t(lambda: ctx.forward(np.arange(n, dtype=np.int32))), flush=True)
    print("eval offset", n, t(lambda: ctx.eval(1000, 1000 + n)), flush=True)
