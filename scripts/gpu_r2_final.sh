#!/bin/bash
# Evidence run of the final build: full GPU suite, smoke, bench lines (default / reference arm / fp32 KV / persistent kernel /
# edit), config 1, EnCodec timings, ncu launch list + full captures of the decode-step kernels and of the codec kernels.
mkdir -p gpurun_out/final3
cd $GRAFT_REPO_ROOT
O=gpurun_out/final3
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,temperature.gpu,power.draw --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu.log 2>&1; echo "exit $?" >> $O/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_steps20.json 2> $O/bench_steps20.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --kv fp32 --no-cpu > $O/bench_kvfp32.json 2> $O/bench_kvfp32.err
VCB_MEGA=1 timeout 600 python bench.py --no-cpu > $O/bench_mega.json 2> $O/bench_mega.err
timeout 600 python bench.py --workload edit --no-cpu > $O/bench_edit_1gpu.json 2> $O/bench_edit_1gpu.err
timeout 300 python scripts/bench_config1.py > $O/config1.json 2> $O/config1.err
for b in 256 32 1; do timeout 600 python scripts/bench_codec.py $b > $O/codec_b$b.json 2>> $O/codec.err; done
VCB_CODEC_PROFILE=1 timeout 600 python scripts/bench_codec.py 256 > /dev/null 2> $O/codec_b256_layers.err
VCB_CODEC_TC=0 timeout 600 python scripts/bench_codec.py 32 > $O/codec_b32_cudacore.json 2>> $O/codec.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python scripts/prof_decode.py 300 2 > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_w_xT|attn_rows|sampler|step_prep" -c 12 -f -o $O/prof_step python scripts/prof_decode.py 300 1 > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc|diag_sum" --launch-skip 203 -c 14 -f -o $O/codec_up python scripts/prof_codec.py 64 100 > $O/ncu_codec.log 2>&1
tail -3 $O/tests_gpu.log; tail -1 $O/smoke.log; head -c 400 $O/bench_default.json; echo; cat $O/codec_b256.json
