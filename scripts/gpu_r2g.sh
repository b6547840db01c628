#!/bin/bash
mkdir -p gpurun_out/r2g
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:mega_step -c 1 -f -o $O/mega python scripts/prof_decode.py 300 1 > $O/ncu_mega.log 2>&1
tail -5 $O/ncu_mega.log
B="python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e"
timeout 300 $B > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python scripts/mega_timeline.py 300 > $O/timeline.txt 2>&1
python -c "
import json
d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]); print(d['ms_per_step'])"
