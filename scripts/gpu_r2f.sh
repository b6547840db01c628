#!/bin/bash
mkdir -p gpurun_out/r2f
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2f
timeout 300 python scripts/mega_timeline.py 300 > $O/timeline.txt 2>&1
VCB_MEGA_NS=6 VCB_MEGA_NB=6 timeout 300 python scripts/mega_timeline.py 300 > $O/timeline_ns6.txt 2>&1
VCB_MEGA=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e > $O/bench_old.json 2> $O/bench_old.err
tail -3 $O/timeline.txt
