#!/bin/bash
# same-box A/B: LayerNorm statistics fetched by the epilogue warps while the mainloop runs (in-tree build) vs the previous
# epilogue order (libvcb200_base.so); then the whole GPU suite on the in-tree build
mkdir -p gpurun_out/r2z
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2z
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu --no-e2e > $O/bench_$name.json 2>> $O/err.txt
  python - <<PY
import json
j=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", j["ms_per_step"], {k[:4]:round(x['ms_per_step'],4) for k,x in j['roofline']['by_kernel'].items()})
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tts_topk40 or batch3 or edit2 or per_utterance or config1 or full_size" > $O/tests_small.log 2>&1; echo "exit $?" >> $O/tests_small.log; tail -3 $O/tests_small.log
run base_1 VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_base.so
run new_1 X=1
run base_2 VCB_LIB=$GRAFT_REPO_ROOT/voicecraft_b200/libvcb200_base.so
run new_2 X=1
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "exit $?" >> $O/tests_gpu.log; tail -3 $O/tests_gpu.log
