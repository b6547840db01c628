#!/bin/bash
mkdir -p gpurun_out
for B in 1 8 16 64; do
  timeout 400 python bench.py --batch $B --steps 300 --no-cpu --no-e2e > gpurun_out/bs_$B.json 2> gpurun_out/bs_$B.err
done
python - <<'PY'
import json
for B in (1, 8, 16, 64):
    try:
        d = json.load(open(f'gpurun_out/bs_{B}.json')); print(B, round(d['value']), round(d['ms_per_step'], 3), round(d['rtf_per_stream'], 1), round(d['step_roofline']['frac'], 3))
    except Exception as e: print(B, 'ERR', open(f'gpurun_out/bs_{B}.err').read()[-300:])
PY
