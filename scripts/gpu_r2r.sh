#!/bin/bash
# compute-sanitizer memcheck over the round-2 code paths (small shapes): device RNG sampler, per-utterance streams, continuous
# batching, persistent kernel, wide prefill, config-1 (head_dim 64), EnCodec tensor-core decoder + encoder
mkdir -p gpurun_out/r2r
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2r
SEL="tts_topk40 or edit2 or batch3 or per_utterance or continuous or device_exponential or generator_stream or persistent_kernel_matches or persistent_kernel_rows or wide_prefill_matches or capacity"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $O/sanitizer_lm.log 2>&1
tail -6 $O/sanitizer_lm.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_codec.py -m gpu -q -k "not full_size" > $O/sanitizer_codec.log 2>&1
tail -6 $O/sanitizer_codec.log
