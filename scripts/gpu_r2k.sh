#!/bin/bash
# ncu of the up-sampling stages of the tensor-core EnCodec decoder (B=64, T=100: skip conv_in + 2 x (ih + 100 steps) = 203 launches)
mkdir -p gpurun_out/r2k
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2k
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc --launch-skip 203 -c 13 -f -o $O/codec_up python scripts/prof_codec.py 64 100 > $O/ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:conv_tc --launch-skip 50 -c 3 -f -o $O/codec_lstm python scripts/prof_codec.py 256 100 > $O/ncu_lstm.log 2>&1
tail -3 $O/ncu.log $O/ncu_lstm.log
