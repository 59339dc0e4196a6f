#!/bin/bash
# Round-2 GPU session G (ONE GPU): records {W,g}, per-warp accumulator pushes, prefetched streaming groups.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2g_tests.txt 2>&1; echo "rc=$?"; tail -8 $O/r2g_tests.txt
echo "== persistent kernel"
for opt in 2 0; do
  TIMELINE_DUMP=$O/r2g_tl_opt${opt}_b256.npy DSGD_PERSIST_OPT=$opt timeout 120 python tools/timeline.py 256 > $O/r2g_timeline_opt${opt}_b256.txt 2>&1; echo "opt=$opt rc=$?"; cat $O/r2g_timeline_opt${opt}_b256.txt
done
for b in 64 1024; do DSGD_PERSIST_OPT=2 timeout 120 python tools/timeline.py $b > $O/r2g_timeline_opt2_b$b.txt 2>&1; head -1 $O/r2g_timeline_opt2_b$b.txt; done
echo "== streaming kernel"
timeout 200 python tools/stream_bw.py > $O/r2g_stream.txt 2>&1; cat $O/r2g_stream.txt
DSGD_STREAM_HOT=1 timeout 200 python tools/stream_bw.py > $O/r2g_stream_hot.txt 2>&1; echo "-- hot"; cat $O/r2g_stream_hot.txt
echo "== ncu: streaming eval + forward(random rows)"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb0 -s 1 -c 1 \
    -o $O/r2g_prof_stream_eval python tools/stream_bw.py > $O/r2g_ncu_se.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb1 -s 8 -c 1 \
    -o $O/r2g_prof_stream_fwd python tools/stream_bw.py > $O/r2g_ncu_sf.log 2>&1; echo "rc=$?"
