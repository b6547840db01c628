#!/bin/bash
mkdir -p gpurun_out/r2n
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n
timeout 900 python scripts/exp_dual_chain.py 300 2 > $O/dual_chain.json 2> $O/dual_chain.err
cat $O/dual_chain.json; tail -5 $O/dual_chain.err
timeout 600 python scripts/bench_codec.py 256 > $O/codec_b256.json 2> $O/codec_b256.err
timeout 600 python scripts/bench_codec.py 32 > $O/codec_b32.json 2>> $O/codec_b256.err
timeout 600 python scripts/bench_codec.py 1 > $O/codec_b1.json 2>> $O/codec_b256.err
cat $O/codec_b256.json $O/codec_b32.json $O/codec_b1.json
