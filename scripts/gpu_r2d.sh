#!/bin/bash
mkdir -p gpurun_out/r2d
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2d
timeout 1200 python -m pytest tests -m gpu -q -x > $O/tests.log 2>&1; echo "exit $?" >> $O/tests.log
timeout 300 python scripts/mega_timeline.py 300 > $O/timeline.txt 2>&1
B="python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err
VCB_MEGA_PF=16 timeout 300 $B > $O/bench_pf16.json 2> $O/bench_pf16.err
VCB_MEGA_NS=12 VCB_MEGA_NB=4 timeout 300 $B > $O/bench_ns12_nb4.json 2> $O/bench_ns12.err
tail -3 $O/tests.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['frac'])" 2>&1 | tail -1; done
