#!/bin/bash
# Round-2 multi-GPU opener: one `gpurun --gpus N` call (N = 2, 4 or 8) that runs the 2-GPU parity tests and the sync bench
# for both exchange schemes of the fused kernel, to separate NVLink BYTES from LATENCY in the step time:
#   DSGD_P2P_MODE=3 (default): weights + gradients as LL words, 16 B per column per peer, no fence
#   DSGD_P2P_MODE=2          : plain fp64, 8 B per column per peer, one system-scope fence + flag per CTA and step
#   DSGD_P2P_MODE=4          : EXPERIMENTAL column ownership (reduce-scatter + all-gather inside the kernel): 2*(K-1)/K*756 KB
#                              per rank and step at any K; written blind at the end of round 1 -- check the parity lines first
# At K GPUs every rank sends its whole dense gradient to every peer: (K-1)*(dim+1)*16 B = 5.3 MB per step at K=8 in mode 3,
# i.e. >= 5.9 us of the 21.3 us step at 900 GB/s per direction.
#   gpurun --gpus 8 --timeout 1500 -- 'bash tools/r2_multi.sh 8'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
for mode in 3 2 4; do
  DSGD_P2P_MODE=$mode timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_multi_tests_mode$mode.txt 2>&1
  echo "multi tests mode=$mode rc=$?"; tail -2 gpurun_out/r2_multi_tests_mode$mode.txt
done
for n in 2 4 8; do
  [ "$n" -le "$N" ] || continue
  for mode in 3 2 4; do
    DSGD_P2P_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n * 10 + mode)) \
        bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_multi_n${n}_mode${mode}.json 2> gpurun_out/r2_multi_n${n}_mode${mode}.err
    python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_multi_n${n}_mode${mode}.json").read().strip().splitlines()[-1])
    print("n=$n mode=$mode value=%.4g ms_per_step=%.5f sgd_steps=%s" % (j["value"], j["ms_per_step"], j["config"].get("sgd_steps_per_bench_step")))
except Exception as e:
    print("n=$n mode=$mode unreadable:", e)
PY
  done
done
