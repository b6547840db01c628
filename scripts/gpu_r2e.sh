#!/bin/bash
mkdir -p gpurun_out/r2e
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2e
timeout 600 python -m pytest tests -m gpu -q -x -k "headline or per_utterance or continuous or chain_path or tokens_match_reference_fixture" > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 300 python scripts/mega_timeline.py 300 > $O/timeline_f5.txt 2>&1
B="python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e"
for cfg in "5 0" "3 0" "8 0" "16 0" "3 16" "4 32" "3 48"; do
  set -- $cfg
  VCB_MEGA_FLIGHT=$1 VCB_MEGA_PF=$2 timeout 300 $B > $O/bench_f$1_pf$2.json 2> $O/bench_f$1_pf$2.err
done
VCB_MEGA_FLIGHT=3 VCB_MEGA_PF=16 timeout 300 python scripts/mega_timeline.py 300 > $O/timeline_f3_pf16.txt 2>&1
tail -3 $O/tests.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['frac'])" 2>&1 | tail -1; done
