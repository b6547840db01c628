#!/bin/bash
mkdir -p gpurun_out/r2s
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2s
SEL="tts_topk40 or edit2 or batch3 or per_utterance or continuous or device_exponential or generator_stream or persistent_kernel_matches or persistent_kernel_rows or wide_prefill_matches or capacity"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SEL" > $O/sanitizer_lm.log 2>&1
grep -c "Invalid\|out of bounds" $O/sanitizer_lm.log; tail -5 $O/sanitizer_lm.log
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "exit $?" >> $O/tests_gpu.log; tail -3 $O/tests_gpu.log
timeout 900 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err; head -c 300 $O/bench_default.json
