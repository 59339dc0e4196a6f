#!/bin/bash
# What the driver runs at round end, plus one ncu capture of the async batch-1 worker:
#   gpurun --timeout 900 -- 'bash tools/gpu_session_final.sh TAG'
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2z}
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1; echo "rc=$?"; tail -2 $O/${TAG}_smoke.txt
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $O/${TAG}_tests.txt 2>&1; echo "rc=$?"; tail -3 $O/${TAG}_tests.txt
echo "== bench (default flags)"; timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "rc=$?"
python - <<PY
import json
j = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g us/step %.3f frac %.4f traffic %s launches %s" % (j["value"], j["e2e"]["value"], j["us_per_sgd_step"], j["roofline"]["frac"], j["roofline"]["traffic"], j["gpu_launches"]))
print("clocks", j["clocks"]); print("cpu_baseline", json.dumps(j["cpu_baseline"])[:700])
print("streaming", {k: (round(v["frac"], 4), v.get("traffic")) for k, v in j["roofline_streaming"].items()})
print("e2e_fit", j["e2e_fit"]["value"], "parity", j["parity"]["max_rel_err_loss"], j["parity"]["max_rel_err_weights"])
PY
echo "== reference arm (short)"; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/${TAG}_bench_ref.json 2>&1; echo "rc=$?"; tail -c 900 $O/${TAG}_bench_ref.json; echo
echo "== ncu --set full: async batch-1 worker (256 lanes)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_async_worker_b1 -c 1 -o $O/${TAG}_prof_async_b1 \
    python bench.py --mode async --lanes 256 --async-updates 200000 --steps 1 --warmup 1 --cpu-seconds 1 > $O/${TAG}_ncu_a.log 2>&1; echo "rc=$?"
