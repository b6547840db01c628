#!/bin/bash
mkdir -p gpurun_out/r2o
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2o
for w in 0 1; do
  VCB_CODEC_LSTM_WIDE=$w VCB_CODEC_PROFILE=1 timeout 600 python scripts/bench_codec.py 256 > $O/codec_b256_wide$w.json 2> $O/codec_b256_wide$w.err
  echo "wide=$w"; grep lstm $O/codec_b256_wide$w.err | tail -4
done
timeout 900 python -m pytest tests/test_codec.py -m gpu -q > $O/tests_codec.log 2>&1; echo "exit $?" >> $O/tests_codec.log
tail -3 $O/tests_codec.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_codec.py -m gpu -q -k "tensor_core or chunking or cuda_decode_matches_fixture" > $O/sanitizer_codec.log 2>&1
tail -12 $O/sanitizer_codec.log
