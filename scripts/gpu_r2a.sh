#!/bin/bash
# first GPU session of round 2: old path (VCB_MEGA=0) with the new RNG / tests, then the persistent kernel
mkdir -p gpurun_out/r2a
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/gpu.txt
export VCB_MEGA=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/tests_mega0.log 2>&1
echo "exit $?" >> gpurun_out/r2a/tests_mega0.log
export VCB_MEGA=1
timeout 900 python -m pytest tests -m gpu -q -x -k "tokens_match_reference_fixture or oracle_kv_bf16 or chain_path or batched_sessions" > gpurun_out/r2a/tests_mega1_small.log 2>&1
echo "exit $?" >> gpurun_out/r2a/tests_mega1_small.log
timeout 900 python -m pytest tests -m gpu -q -k "headline or full_size or per_utterance or generator_stream" > gpurun_out/r2a/tests_mega1_big.log 2>&1
echo "exit $?" >> gpurun_out/r2a/tests_mega1_big.log
timeout 600 python bench.py --steps 600 --warmup 10 --no-cpu > gpurun_out/r2a/bench_mega1.json 2> gpurun_out/r2a/bench_mega1.err
VCB_MEGA=0 timeout 600 python bench.py --steps 600 --warmup 10 --no-cpu --no-e2e > gpurun_out/r2a/bench_mega0.json 2> gpurun_out/r2a/bench_mega0.err
tail -3 gpurun_out/r2a/*.log; cat gpurun_out/r2a/bench_mega1.json | head -c 1500
