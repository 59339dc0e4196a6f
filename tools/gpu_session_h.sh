#!/bin/bash
# Round-2 GPU session H (ONE GPU): rolling indices, acquire-fence experiment (DSGD_PERSIST_OPT bit 0), stream kernel at 1024 threads.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2h_tests.txt 2>&1; echo "rc=$?"; tail -5 $O/r2h_tests.txt
for opt in 2 3; do
  DSGD_PERSIST_OPT=$opt timeout 120 python tools/timeline.py 256 > $O/r2h_timeline_opt${opt}_b256.txt 2>&1; echo "opt=$opt rc=$?"; grep -v "^step 10[123]" $O/r2h_timeline_opt${opt}_b256.txt
  DSGD_PERSIST_OPT=$opt timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_one_gpu.py -m gpu -q -p no:cacheprovider > $O/r2h_parity_opt$opt.txt 2>&1; echo "parity opt=$opt rc=$?"; tail -2 $O/r2h_parity_opt$opt.txt
done
echo "== streaming kernel"
timeout 200 python tools/stream_bw.py > $O/r2h_stream.txt 2>&1; cat $O/r2h_stream.txt
echo "== bench (no extras)"
timeout 400 python bench.py --no-extras --cpu-seconds 3 > $O/r2h_bench.json 2> $O/r2h_bench.err; echo "rc=$?"; python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2h_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4g us/step %.3f frac %.4f" % (j["value"], j["e2e"]["value"], j["us_per_sgd_step"], j["roofline"]["frac"]))
print(json.dumps(j["roofline_streaming"])[:900])
PY
