#!/bin/bash
# Round-2 opener: ONE gpurun call that (1) re-runs the GPU suite, (2) runs the latency microbenchmarks behind the design
# questions of DESIGN.md section 8, (3) checks parity and measures the bench line for every experimental variant of the
# persistent kernel (DSGD_PERSIST_OPT, dsgd_persistent.cuh kOpt), (4) prints CTA 0's step timeline for each and (5) does
# the same parity + bandwidth check for the streaming-pass variants (DSGD_STREAM_OPT, dsgd_stream_x.cuh).
#   gpurun --timeout 2400 -- 'bash tools/r2_first_call.sh'            # everything, ~35 GPU-minutes
#   gpurun --timeout 900  -- 'bash tools/r2_first_call.sh micro persist'   # sections: tests micro persist stream async
# Everything lands in gpurun_out/r2_*.txt|json.  Nothing here is a bench value of record (bench.py alone is).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SECTIONS="${*:-tests micro persist stream async}"
want() { case " $SECTIONS " in *" $1 "*) return 0;; *) return 1;; esac; }
[ -x tools/microbench ] || (cd tools && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu)
if want tests; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.txt 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_tests.txt; fi
if want micro; then timeout 300 ./tools/microbench > gpurun_out/r2_microbench.txt 2>&1; echo "microbench rc=$?"; tail -25 gpurun_out/r2_microbench.txt; fi
for opt in 0 1 2 4 7; do   # bit 0 flag barrier, bit 1 pushed partials, bit 2 one-pass single-chunk rows; 7 = all
  want persist || break
  export DSGD_PERSIST_OPT=$opt
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfullsize.py -q -m gpu -k "trajectory or epoch or golden or overflow" \
      > gpurun_out/r2_parity_opt$opt.txt 2>&1; echo "parity opt=$opt rc=$?"
  timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_opt$opt.json 2> gpurun_out/r2_bench_opt$opt.err; echo "bench opt=$opt rc=$?"
  timeout 120 python tools/timeline.py 256 > gpurun_out/r2_timeline_opt$opt.txt 2>&1
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_bench_opt$opt.json").read().strip().splitlines()[-1])
    print("opt=$opt value=%.4g e2e=%.4g ms_per_step=%.5f frac=%.4f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["roofline"]["frac"]))
except Exception as e:
    print("opt=$opt bench line unreadable:", e)
PY
done
unset DSGD_PERSIST_OPT
# streaming pass variants (dsgd_stream_x.cuh): 1 = fp32 fast path, 2 = hot-column accumulators (scatter), 3 = both
for opt in 0 1 2 3; do
  want stream || break
  export DSGD_STREAM_OPT=$opt
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfullsize.py -q -m gpu \
      -k "streaming or eval or gradient or zero_weights or additive or shards" > gpurun_out/r2_stream_parity_opt$opt.txt 2>&1
  echo "stream parity opt=$opt rc=$?"
  timeout 300 python tools/stream_bw.py > gpurun_out/r2_stream_bw_opt$opt.txt 2>&1; echo "--- DSGD_STREAM_OPT=$opt"; cat gpurun_out/r2_stream_bw_opt$opt.txt
done
unset DSGD_STREAM_OPT
# async batch-1 fast path (k_async_worker_b1, DSGD_ASYNC_OPT=1)
for opt in 0 1; do
  want async || break
  export DSGD_ASYNC_OPT=$opt
  timeout 300 python -m pytest tests/test_gpu_async.py -q -m gpu > gpurun_out/r2_async_parity_opt$opt.txt 2>&1; echo "async parity opt=$opt rc=$?"
  timeout 300 python bench.py --mode async --steps 5 --warmup 3 > gpurun_out/r2_bench_async_opt$opt.json 2> gpurun_out/r2_bench_async_opt$opt.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_bench_async_opt$opt.json").read().strip().splitlines()[-1])
    print("async opt=$opt value=%.4g e2e=%.4g" % (j["value"], j["e2e"]["value"]))
except Exception as e:
    print("async opt=$opt bench line unreadable:", e)
PY
done
unset DSGD_ASYNC_OPT
