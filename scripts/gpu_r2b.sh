#!/bin/bash
# GPU session 2 of round 2: full suite on the persistent kernel, its per-phase timeline, ring / prefetch sweeps
mkdir -p gpurun_out/r2b
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b
timeout 1200 python -m pytest tests -m gpu -q > $O/tests_mega1.log 2>&1; echo "exit $?" >> $O/tests_mega1.log
VCB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -k experimental > $O/tests_grouped_prefill.log 2>&1; echo "exit $?" >> $O/tests_grouped_prefill.log
timeout 300 python scripts/mega_timeline.py 300 > $O/timeline_default.txt 2>&1
B="python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err
VCB_MEGA_PF=16 timeout 300 $B > $O/bench_pf16.json 2> $O/bench_pf16.err
VCB_MEGA_PF=48 timeout 300 $B > $O/bench_pf48.json 2> $O/bench_pf48.err
VCB_MEGA_NS=12 VCB_MEGA_NB=4 timeout 300 $B > $O/bench_ns12_nb4.json 2> $O/bench_ns12.err
VCB_MEGA_NS=10 VCB_MEGA_NB=8 timeout 300 $B > $O/bench_ns10_nb8.json 2> $O/bench_ns10.err
VCB_MEGA_GRID=132 timeout 300 $B > $O/bench_grid132.json 2> $O/bench_grid132.err
VCB_PREFILL_ATT_GROUP=4 timeout 300 python scripts/prof_prefill.py > $O/prefill_group4.txt 2>&1
timeout 300 python scripts/prof_prefill.py > $O/prefill_default.txt 2>&1
tail -4 $O/tests_mega1.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['frac'])" 2>&1 | tail -1; done
