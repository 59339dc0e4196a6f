#!/bin/bash
# Round-2 GPU session D (ONE GPU): fused K-rank tests on one GPU (with dsgd_reserve), streaming kernel with pipelined claims.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== fused K-rank tests on one GPU"; timeout 600 python -m pytest tests/test_gpu_fused_one_gpu.py -m gpu -q -p no:cacheprovider > $O/r2d_tests_fused.txt 2>&1; echo "rc=$?"; tail -15 $O/r2d_tests_fused.txt
echo "== other gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fused_one_gpu.py > $O/r2d_tests.txt 2>&1; echo "rc=$?"; tail -8 $O/r2d_tests.txt
echo "== streaming kernel"
timeout 200 python tools/stream_bw.py > $O/r2d_stream_flat.txt 2>&1; cat $O/r2d_stream_flat.txt
echo "== ncu --set full: streaming forward over 262144 random rows, eval"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb1 -s 8 -c 1 \
    -o $O/r2d_prof_stream_fwd python tools/stream_bw.py > $O/r2d_ncu_sf.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_stream_rowsILb0ELb0 -s 1 -c 1 \
    -o $O/r2d_prof_stream_eval python tools/stream_bw.py > $O/r2d_ncu_se.log 2>&1; echo "rc=$?"
