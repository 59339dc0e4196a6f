#!/bin/bash
# Streaming-kernel check: parity tests of the streaming paths, then tools/stream_bw.py.
#   gpurun --timeout 600 -- 'bash tools/gpu_session_stream.sh TAG'
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r2s}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfullsize.py -m gpu -q -p no:cacheprovider > $O/${TAG}_tests.txt 2>&1; echo "tests rc=$?"; tail -3 $O/${TAG}_tests.txt
timeout 200 python tools/stream_bw.py > $O/${TAG}_stream.txt 2>&1; cat $O/${TAG}_stream.txt
