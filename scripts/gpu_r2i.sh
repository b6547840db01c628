#!/bin/bash
# first hardware run of the tensor-core EnCodec decoder: stage-by-stage check, codec tests, timing
mkdir -p gpurun_out/r2j
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2j
timeout 300 python scripts/codec_tc_debug.py 16 3 37 0 > $O/dbg_nf16_nolstm.txt 2>&1
timeout 300 python scripts/codec_tc_debug.py 16 3 37 2 > $O/dbg_nf16_lstm2.txt 2>&1
timeout 300 python scripts/codec_tc_debug.py 8 2 24 2 > $O/dbg_nf8.txt 2>&1
timeout 600 python scripts/codec_tc_debug.py 64 2 53 2 > $O/dbg_full.txt 2>&1
timeout 900 python -m pytest tests/test_codec.py -m gpu -q > $O/tests_codec.log 2>&1; echo "exit $?" >> $O/tests_codec.log
VCB_CODEC_PROFILE=1 timeout 600 python scripts/bench_codec.py 32 > $O/codec_b32.json 2> $O/codec_b32.err
VCB_CODEC_PROFILE=1 timeout 900 python scripts/bench_codec.py 256 > $O/codec_b256.json 2> $O/codec_b256.err
for f in dbg_nf16_nolstm dbg_nf16_lstm2 dbg_nf8 dbg_full; do echo "== $f"; tail -22 $O/$f.txt; done
tail -5 $O/tests_codec.log; cat $O/codec_b256.json; tail -30 $O/codec_b256.err
