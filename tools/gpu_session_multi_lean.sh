#!/bin/bash
# Lean multi-GPU check: `gpurun --gpus N -- 'bash tools/gpu_session_multi_lean.sh N TAG'`: the 2-GPU fused parity test, the
# step timeline and a short bench line (no sub-records) at N GPUs.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
TAG=${2:-r2m}
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fused_one_gpu.py -m gpu -q -p no:cacheprovider -k "p2p or fused" > $O/${TAG}_tests.txt 2>&1; echo "fused 2-GPU test rc=$?"; tail -2 $O/${TAG}_tests.txt
timeout 200 python tools/timeline.py 256 --world $N > $O/${TAG}_timeline_n$N.txt 2>&1; echo "timeline n=$N rc=$?"; grep -v "^step 10[123]" $O/${TAG}_timeline_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --no-extras --steps 3 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "bench n=$N rc=$?"
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("n=$N value %.4g e2e %.4g us/step %.3f" % (j["value"], j["e2e"]["value"], j["us_per_sgd_step"]))
except Exception as e:
    print("n=$N bench line unreadable:", e)
PY
