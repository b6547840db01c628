#!/bin/bash
# same-box A/B: default build vs a build with the device-timeline marks compiled out (-DVCB_NO_TIMELINE)
mkdir -p gpurun_out/r2v
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2v
for rep in 1 2; do
  for v in cur notl; do
    if [ $v = notl ]; then export VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_notl.so; else unset VCB_LIB; fi
    timeout 600 python bench.py --no-cpu --no-e2e > $O/bench_${v}_$rep.json 2>> $O/err.txt
    python - <<PY
import json
j=json.loads(open("$O/bench_${v}_$rep.json").read().strip().splitlines()[-1]); print("$v $rep", j["ms_per_step"], {k[:4]:round(x['ms_per_step'],4) for k,x in j['roofline']['by_kernel'].items()})
PY
  done
done
