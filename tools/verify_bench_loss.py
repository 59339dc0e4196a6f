"""Post-hoc parity check of a multi-GPU bench run: recompute the whole trajectory of `bench.py --gpus N` (same seeds,
same batch draws) with the fp64 CPU oracle and compare the last step's loss with the `final_batch_loss` on the JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from distributed_sgd_b200.utils import synthetic_rcv1
from oracle.oracle import Oracle

line = json.load(open(sys.argv[1]))
K, steps, warm = line["n_gpus"], line["steps"], line["warmup"]
B, S = line["config"]["batch_per_gpu"], line["config"]["sgd_steps_per_bench_step"]
data = synthetic_rcv1(n_rows=bench.N_ROWS, dim=bench.DIM, seed=0)
n_train = int(data.n_rows * bench.TRAIN_FRAC)
orc = Oracle(data.row_ptr, data.col, data.val, data.label, data.dim, bench.LAMBDA)
orc.set_dim_sparsity(orc.dim_sparsity(n_train))
per = n_train // K
total = warm + steps
parts = []
for rank in range(K):
    rng = np.random.default_rng(0 * 1000 + rank)
    parts.append(bench.draw_batches(rng, rank * per, (rank + 1) * per, B, S * total).reshape(total * S, B))
idx = np.stack(parts, axis=1).reshape(-1)            # step-major, then worker-major
t = time.time()
w, losses = orc.sync_steps(np.zeros(data.dim), idx, [B] * K, bench.LR, n_steps=total * S, threads=K)
print(f"oracle: {total*S} steps x {K} workers in {time.time()-t:.1f} s; last-step loss {losses[-1]!r}; bench line {line['final_batch_loss']!r}; "
      f"rel diff {abs(losses[-1]-line['final_batch_loss'])/abs(losses[-1]):.3e}")
if "final_weights_l1" in line:
    l1 = float(np.abs(w).sum())
    print(f"final weights: oracle l1 {l1!r} nnz {int(np.count_nonzero(w))}; bench l1 {line['final_weights_l1']!r} nnz {line['final_weights_nnz']}; "
          f"rel diff {abs(l1-line['final_weights_l1'])/l1:.3e}")
