#!/bin/bash
# same-box A/B: default build (timeline marks compiled out) vs v2 (L2 prefetch of the next GEMM's weights issued after the
# cluster barrier) vs the prefetch knob settings
mkdir -p gpurun_out/r2w
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2w
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-e2e > $O/bench_$name.json 2>> $O/err.txt
  python - <<PY
import json
j=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", j["ms_per_step"], {k[:4]:round(x['ms_per_step'],4) for k,x in j['roofline']['by_kernel'].items()})
PY
}
run cur_1 X=1
run v2_1 VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_v2.so
run pf2_1 VCB_PREFETCH=2
run pf0_1 VCB_PREFETCH=0
run cur_2 X=1
run v2_2 VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_v2.so
