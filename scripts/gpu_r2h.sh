#!/bin/bash
mkdir -p gpurun_out/r2h
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h
export VCB_MEGA=1
B="python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e"
for cfg in "5 11 6" "8 11 6" "11 11 6" "12 12 4" "8 10 8"; do
  set -- $cfg
  VCB_MEGA_FLIGHT=$1 VCB_MEGA_NS=$2 VCB_MEGA_NB=$3 timeout 300 $B > $O/bench_f$1_ns$2_nb$3.json 2> $O/bench_f$1_ns$2_nb$3.err
done
VCB_MEGA_FLIGHT=11 timeout 300 python scripts/mega_timeline.py 300 > $O/timeline_f11.txt 2>&1
for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['frac'])" 2>&1 | tail -1; done
