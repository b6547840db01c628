#!/bin/bash
mkdir -p gpurun_out/$1
cd $GRAFT_REPO_ROOT
O=gpurun_out/$1
timeout 600 python scripts/codec_tc_debug.py 64 2 53 2 > $O/dbg_full.txt 2>&1
timeout 900 python -m pytest tests/test_codec.py -m gpu -q > $O/tests_codec.log 2>&1; echo "exit $?" >> $O/tests_codec.log
VCB_CODEC_PROFILE=1 timeout 600 python scripts/bench_codec.py 32 > $O/codec_b32.json 2> $O/codec_b32.err
VCB_CODEC_PROFILE=1 timeout 900 python scripts/bench_codec.py 256 > $O/codec_b256.json 2> $O/codec_b256.err
grep waveform $O/dbg_full.txt; tail -4 $O/tests_codec.log; cat $O/codec_b32.json $O/codec_b256.json; grep codec_tc $O/codec_b256.err | tail -17
