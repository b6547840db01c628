#!/bin/bash
# bench lines + ncu of the decode-step kernels for the final build (the GPU suite of this build: gpurun_out/r2y, 86 passed)
mkdir -p gpurun_out/final4
cd $GRAFT_REPO_ROOT
O=gpurun_out/final4
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_steps20.json 2> $O/bench_steps20.err
timeout 600 python bench.py --kv fp32 --no-cpu > $O/bench_kvfp32.json 2> $O/bench_kvfp32.err
timeout 600 python bench.py --workload edit --no-cpu > $O/bench_edit_1gpu.json 2> $O/bench_edit_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python scripts/prof_decode.py 300 2 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_w_xT|attn_rows|sampler|step_prep" -c 12 -f -o $O/prof_step python scripts/prof_decode.py 300 1 > $O/ncu_full.log 2>&1
head -c 300 $O/bench_default.json; echo
