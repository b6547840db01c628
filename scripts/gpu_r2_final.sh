#!/bin/bash
# Evidence run of the default (per-kernel) decode path: full GPU suite, bench lines, ncu launch list + full captures.
mkdir -p gpurun_out/final
cd $GRAFT_REPO_ROOT
O=gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "exit $?" >> $O/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --kv fp32 --no-cpu > $O/bench_kvfp32.json 2> $O/bench_kvfp32.err
VCB_MEGA=1 timeout 600 python bench.py --no-cpu > $O/bench_mega.json 2> $O/bench_mega.err
timeout 600 python bench.py --workload edit --no-cpu > $O/bench_edit_1gpu.json 2> $O/bench_edit_1gpu.err
timeout 300 python scripts/bench_config1.py > $O/config1.json 2> $O/config1.err
timeout 600 python scripts/bench_codec.py 256 > $O/codec_b256.json 2> $O/codec.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python scripts/prof_decode.py 300 2 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_w_xT|attn_rows|sampler|step_prep" -c 12 -f -o $O/prof_step python scripts/prof_decode.py 300 1 > $O/ncu_full.log 2>&1
tail -3 $O/tests_gpu.log; cat $O/smoke.log | tail -1; head -c 600 $O/bench_default.json
