#!/bin/bash
mkdir -p gpurun_out/r2u
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2u
for i in 1 2; do timeout 600 python bench.py --no-cpu --no-e2e > $O/bench_$i.json 2>> $O/err.txt; done
for f in $O/bench_*.json; do python - <<PY
import json
j=json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", j["ms_per_step"], {k:round(v['ms_per_step'],4) for k,v in j['roofline']['by_kernel'].items()})
PY
done
nvidia-smi --query-gpu=name,uuid,clocks.max.sm,clocks.max.mem,power.limit --format=csv
