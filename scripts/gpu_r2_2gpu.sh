#!/bin/bash
# 2-GPU evidence (gpurun --gpus 2): tts (different utterances per rank + NCCL gather in e2e), edit (16 utterances
# partitioned, result checked against a single-GPU decode), reference arm under torchrun (rank 0 only).
mkdir -p gpurun_out/g2
cd $GRAFT_REPO_ROOT
O=gpurun_out/g2
nvidia-smi -L > $O/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --no-cpu > $O/bench_tts_2gpu.json 2> $O/bench_tts_2gpu.err
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --workload edit --no-cpu > $O/bench_edit_2gpu.json 2> $O/bench_edit_2gpu.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --impl reference --steps 4 --warmup 1 > $O/bench_ref_2gpu.json 2> $O/bench_ref_2gpu.err
head -c 1500 $O/bench_tts_2gpu.json; echo; cat $O/bench_edit_2gpu.json; tail -n 3 $O/bench_tts_2gpu.err; tail -n 3 $O/bench_edit_2gpu.err
