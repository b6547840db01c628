#!/bin/bash
# GPU session 3: warp-per-page attention + canonical merge + relaxed flag polling
mkdir -p gpurun_out/r2c
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c
timeout 1200 python -m pytest tests -m gpu -q > $O/tests_mega1.log 2>&1; echo "exit $?" >> $O/tests_mega1.log
timeout 300 python scripts/mega_timeline.py 300 > $O/timeline_default.txt 2>&1
B="python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err
VCB_MEGA_PF=16 timeout 300 $B > $O/bench_pf16.json 2> $O/bench_pf16.err
VCB_MEGA_NS=12 VCB_MEGA_NB=4 timeout 300 $B > $O/bench_ns12_nb4.json 2> $O/bench_ns12.err
VCB_MEGA_NS=10 VCB_MEGA_NB=8 timeout 300 $B > $O/bench_ns10_nb8.json 2> $O/bench_ns10.err
VCB_MEGA_NS=8 VCB_MEGA_NB=8 timeout 300 $B > $O/bench_ns8_nb8.json 2> $O/bench_ns8.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --kv fp32 > $O/bench_kvfp32.json 2> $O/bench_kvfp32.err
tail -4 $O/tests_mega1.log; for f in $O/bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['step_roofline']['frac'])" 2>&1 | tail -1; done
