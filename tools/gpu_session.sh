#!/bin/bash
# The one-GPU evidence session of a round:  gpurun --timeout 1800 -- 'bash tools/gpu_session.sh TAG'
# GPU suite, step timelines, streaming bandwidths, the bench line, the ncu launch list of the bench command and one
# `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/TAG_*; tools/ncu_extract.py turns the captures
# into profiles/ncu_traffic.json + a summary.  Nothing printed under ncu is a bench value.
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_tests.txt 2>&1; echo "rc=$?"; tail -4 $O/${TAG}_tests.txt
echo "== step timelines"
for b in 256 64 1024; do timeout 120 python tools/timeline.py $b > $O/${TAG}_timeline_b$b.txt 2>&1; grep -v "^step 10[123]" $O/${TAG}_timeline_b$b.txt; done
echo "== streaming kernel"; timeout 200 python tools/stream_bw.py > $O/${TAG}_stream.txt 2>&1; cat $O/${TAG}_stream.txt
echo "== bench line"
timeout 900 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -3 $O/${TAG}_bench_n1.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g us/step %.3f frac %.4f" % (j["value"], j["e2e"]["value"], j["us_per_sgd_step"], j["roofline"]["frac"]))
    for k in ("sweep", "parity", "e2e_fit", "async", "rpc_seam", "roofline_streaming", "cpu_baseline"):
        print(k, json.dumps(j.get(k))[:900])
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --sgd-steps 500 --cpu-seconds 1 --no-extras > $O/${TAG}_ncu_bench.log 2>&1; echo "launch list rc=$?"
echo "== ncu --set full: persistent kernel (300 steps, batch 256)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sync_persistent -s 1 -c 1 -o $O/${TAG}_prof_persist \
    python tools/timeline.py 256 > $O/${TAG}_ncu_p.log 2>&1; echo "rc=$?"
echo "== ncu --set full: streaming kernels inside the bench (eval over the train rows; gradient of 262144 rows on trained weights)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb0 -s 1 -c 1 \
    -o $O/${TAG}_prof_stream_eval python bench.py --steps 1 --warmup 3 --cpu-seconds 1 --no-extras > $O/${TAG}_ncu_se.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb1ELb0 -s 1 -c 1 \
    -o $O/${TAG}_prof_stream_scatter python bench.py --steps 1 --warmup 3 --cpu-seconds 1 --no-extras > $O/${TAG}_ncu_ss.log 2>&1; echo "rc=$?"
ls -la $O | grep ${TAG}_ | head -40
