#!/bin/bash
# Round-2 multi-GPU session: `gpurun --gpus N -- 'bash tools/gpu_session_multi.sh N'`.  2-GPU parity tests, the step timeline of
# the fused kernel at every GPU count up to N, and the bench line (parity / sweep / async / nvlink sub-records) at N.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
TAG=${2:-r2m}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -8
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_async.py -m gpu -q -p no:cacheprovider > $O/${TAG}_tests.txt 2>&1; echo "rc=$?"; tail -4 $O/${TAG}_tests.txt
for n in 2 4 8; do
  [ "$n" -le "$N" ] || continue
  timeout 200 python tools/timeline.py 256 --world $n > $O/${TAG}_timeline_n$n.txt 2>&1; echo "timeline n=$n rc=$?"; grep -v "^step 10[123]" $O/${TAG}_timeline_n$n.txt
done
for n in 2 4 8; do
  [ "$n" -le "$N" ] || continue
  [ "$n" -eq "$N" ] && EXTRA="" || EXTRA="--no-extras"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n $EXTRA > $O/${TAG}_bench_n$n.json 2> $O/${TAG}_bench_n$n.err; echo "bench n=$n rc=$?"; tail -2 $O/${TAG}_bench_n$n.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_n$n.json").read().strip().splitlines()[-1])
    print("n=$n value %.4g e2e %.4g us/step %.3f" % (j["value"], j["e2e"]["value"], j["us_per_sgd_step"]))
    for k in ("nvlink", "parity", "sweep", "e2e_fit", "async"):
        if j.get(k) is not None: print("  ", k, json.dumps(j[k])[:700])
except Exception as e:
    print("n=$n bench line unreadable:", e)
PY
done
