"""Dev tool: samples/s of the device-resident sync loop for several CTA counts / batch sizes (one GPU).
Each configuration runs in a fresh process because the CTA count is read from the environment once."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, time
sys.path.insert(0, %r)
import numpy as np
from distributed_sgd_b200.native import NativeCtx
from distributed_sgd_b200.utils import synthetic_rcv1
B, S = int(sys.argv[1]), int(sys.argv[2])
data = synthetic_rcv1(n_rows=200000, seed=0)
n_train = 160000
ctx = NativeCtx(0, data.dim, 1e-5)
ctx.load_csr(data.row_ptr, data.col, data.val, data.label)
ctx.compute_dim_sparsity(n_train)
rng = np.random.default_rng(0)
idx = np.stack([rng.choice(n_train, size=B, replace=False) for _ in range(S)]).astype(np.int32).reshape(-1)
ctx.stage_samples(idx)
ctx.set_weights(np.zeros(data.dim))
for _ in range(2):
    ctx.sync_steps_staged(0, B, S, 0.5, want_losses=True)
ctx.synchronize()
best = 1e9
for _ in range(3):
    ctx.timer_start()
    ctx.sync_steps_staged(0, B, S, 0.5, want_losses=True)
    best = min(best, ctx.timer_stop())
print(json.dumps({"batch": B, "steps": S, "us_per_step": best * 1e3 / S, "msamples_per_s": B * S / best / 1e3}))
''' % ROOT

for batch in (64, 256, 1024):
    for g in (os.environ.get("SWEEP_G", "16,32,64,100,148").split(",")):
        env = dict(os.environ, DSGD_PERSIST_CTAS=g)
        r = subprocess.run([sys.executable, "-c", CHILD, str(batch), "2000"], env=env, capture_output=True, text=True)
        print("G=%s" % g, r.stdout.strip() or r.stderr[-400:], flush=True)
