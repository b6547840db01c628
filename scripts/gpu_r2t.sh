#!/bin/bash
# completion-counter hand-over between the kernels of a decode step (VCB_DEPCTR=1) vs griddepcontrol.wait
mkdir -p gpurun_out/r2t
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2t
VCB_DEPCTR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tts_topk40 or batch3 or edit2 or per_utterance or continuous or config1" > $O/tests_dep_small.log 2>&1; echo "exit $?" >> $O/tests_dep_small.log; tail -3 $O/tests_dep_small.log
for v in 0 1 0 1; do
  VCB_DEPCTR=$v timeout 600 python bench.py --no-cpu --no-e2e > $O/bench_dep${v}_$RANDOM.json 2>> $O/err.txt
done
for f in $O/bench_dep*.json; do python - <<PY
import json
j=json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", j["ms_per_step"], j["value"])
PY
done
VCB_DEPCTR=1 timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu_dep1.log 2>&1; echo "exit $?" >> $O/tests_gpu_dep1.log; tail -3 $O/tests_gpu_dep1.log
