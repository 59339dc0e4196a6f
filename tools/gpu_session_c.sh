#!/bin/bash
# Round-2 GPU session C (ONE GPU): all-to-all flag barrier + pushed partials, RED scatter, lean flat-stream kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_zfullsize.py > $O/r2c_tests.txt 2>&1; echo "rc=$?"; tail -12 $O/r2c_tests.txt
echo "== persistent kernel (bit 1 = one-pass single-chunk rows)"
for opt in 0 2; do
  for b in 256; do
    TIMELINE_DUMP=$O/r2c_tl_opt${opt}_b$b.npy DSGD_PERSIST_OPT=$opt timeout 120 python tools/timeline.py $b > $O/r2c_timeline_opt${opt}_b$b.txt 2>&1; echo "opt=$opt batch=$b rc=$?"; cat $O/r2c_timeline_opt${opt}_b$b.txt
  done
done
for b in 64 1024; do DSGD_PERSIST_OPT=2 timeout 120 python tools/timeline.py $b > $O/r2c_timeline_opt2_b$b.txt 2>&1; head -1 $O/r2c_timeline_opt2_b$b.txt; done
echo "== streaming kernel variants"
timeout 200 python tools/stream_bw.py > $O/r2c_stream_flat.txt 2>&1; echo "-- flat stream (lean)"; cat $O/r2c_stream_flat.txt
DSGD_STREAM_HOT=1 timeout 200 python tools/stream_bw.py > $O/r2c_stream_flat_hot.txt 2>&1; echo "-- + hot-column accumulators"; cat $O/r2c_stream_flat_hot.txt
echo "== full-size tests"; timeout 600 python -m pytest tests/test_gpu_zfullsize.py -m gpu -q -p no:cacheprovider > $O/r2c_tests_full.txt 2>&1; echo "rc=$?"; tail -5 $O/r2c_tests_full.txt
echo "== ncu --set full: streaming eval"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb0 -s 1 -c 1 \
    -o $O/r2c_prof_stream_eval python tools/stream_bw.py > $O/r2c_ncu_se.log 2>&1; echo "rc=$?"
echo "== ncu --set full: persistent kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sync_persistent -s 1 -c 1 -o $O/r2c_prof_persist \
    python tools/timeline.py 256 > $O/r2c_ncu_p.log 2>&1; echo "rc=$?"
