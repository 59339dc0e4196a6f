#!/bin/bash
# Round-2 GPU session A (ONE GPU): the GPU suite on the rewritten kernels, A/B of the temporary variant switches, the bench
# line, and the ncu evidence.  Everything lands in gpurun_out/r2a_*.  Nothing printed under ncu is a bench value.
#   gpurun --timeout 1500 -- 'bash tools/gpu_session_a.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2a_tests.txt 2>&1; echo "rc=$?"; tail -15 $O/r2a_tests.txt
echo "== persistent kernel variants (bit 0 flag barrier, bit 1 one-pass single-chunk rows)"
for opt in 0 1 2 3; do
  for b in 256; do
    DSGD_PERSIST_OPT=$opt timeout 120 python tools/timeline.py $b > $O/r2a_timeline_opt${opt}_b$b.txt 2>&1; echo "opt=$opt batch=$b rc=$?"; head -1 $O/r2a_timeline_opt${opt}_b$b.txt
  done
done
for b in 64 1024; do DSGD_PERSIST_OPT=2 timeout 120 python tools/timeline.py $b > $O/r2a_timeline_opt2_b$b.txt 2>&1; head -1 $O/r2a_timeline_opt2_b$b.txt; done
echo "== streaming kernel variants"
DSGD_STREAM_V1=1 timeout 200 python tools/stream_bw.py > $O/r2a_stream_v1.txt 2>&1; echo "-- v1 (round 1)"; cat $O/r2a_stream_v1.txt
timeout 200 python tools/stream_bw.py > $O/r2a_stream_flat.txt 2>&1; echo "-- flat stream"; cat $O/r2a_stream_flat.txt
DSGD_STREAM_HOT=1 timeout 200 python tools/stream_bw.py > $O/r2a_stream_flat_hot.txt 2>&1; echo "-- flat stream + hot-column accumulators"; cat $O/r2a_stream_flat_hot.txt
echo "== bench line"
timeout 600 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench rc=$?"; tail -3 $O/r2a_bench.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2a_bench.json").read().strip().splitlines()[-1])
    print("value %.4g e2e %.4g us/step %.3f frac %.4f" % (j["value"], j["e2e"]["value"], j["us_per_sgd_step"], j["roofline"]["frac"]))
    for k in ("sweep", "parity", "e2e_fit", "async", "rpc_seam", "roofline_streaming"):
        print(k, json.dumps(j.get(k))[:600])
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== ncu launch list of the bench command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r2a_launches.csv \
    python bench.py --steps 2 --warmup 1 --sgd-steps 500 --cpu-seconds 1 --no-extras > $O/r2a_ncu_bench.log 2>&1; echo "launch list rc=$?"
echo "== ncu --set full: persistent kernel (300 steps, batch 256)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sync_persistent -s 1 -c 1 -o $O/r2a_prof_persist \
    python tools/timeline.py 256 > $O/r2a_ncu_p.log 2>&1; echo "rc=$?"
echo "== ncu --set full: streaming kernels (eval over the train rows; gradient of 262144 rows)"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb0 -s 1 -c 1 \
    -o $O/r2a_prof_stream_eval python tools/stream_bw.py > $O/r2a_ncu_se.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb1 -s 3 -c 1 \
    -o $O/r2a_prof_stream_scatter python tools/stream_bw.py > $O/r2a_ncu_ss.log 2>&1; echo "rc=$?"
ls -la $O | grep r2a
