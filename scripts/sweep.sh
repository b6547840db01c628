#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; }
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name $(echo $envs | tr ',' ' ')
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sw_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['ms_per_step'],3), d.get('gemm_chain'), d['gpu_launches'])
    except Exception as e: print(f,'ERR', open(f.replace('.json','.err')).read()[-300:])
PY
