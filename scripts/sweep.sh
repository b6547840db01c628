#!/bin/bash
# usage: scripts/sweep.sh  -- runs bench.py under several env settings (GPU box)
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; }
run base VCB_PDL=1
run max148 VCB_GEMM_MAXCTAS=148
run st3 VCB_GEMM_STAGES=3
run st2 VCB_GEMM_STAGES=2
run max148_st3 VCB_GEMM_MAXCTAS=148 VCB_GEMM_STAGES=3
run max148_st6 VCB_GEMM_MAXCTAS=148 VCB_GEMM_STAGES=6
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sw_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],3))
    except Exception as e: print(f,'ERR')
PY
