#!/bin/bash
# A/B on one box: previous epilogue (libvcb200_prev.so) vs current build, alternating, no profiling
mkdir -p gpurun_out/r2q
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q
for rep in 1 2; do
  for v in prev cur; do
    if [ $v = prev ]; then export VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_prev.so; else unset VCB_LIB; fi
    timeout 600 python scripts/bench_codec.py 256 > $O/b256_${v}_$rep.json 2>> $O/err.txt
    timeout 600 python scripts/bench_codec.py 32 > $O/b32_${v}_$rep.json 2>> $O/err.txt
    echo "$v $rep: $(cut -c60-110 $O/b256_${v}_$rep.json) | $(cut -c58-100 $O/b32_${v}_$rep.json)"
  done
done
for v in prev cur; do
  if [ $v = prev ]; then export VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_prev.so; else unset VCB_LIB; fi
  VCB_CODEC_PROFILE=1 timeout 600 python scripts/bench_codec.py 256 > /dev/null 2> $O/layers_$v.txt
  echo "== $v"; grep codec_tc $O/layers_$v.txt | tail -18 | awk '{printf "%s %s | ", $2, $3} END{print ""}'
done
