#!/bin/bash
mkdir -p gpurun_out/last
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_codec.py -m gpu -q -k "causal_at_full_length" > gpurun_out/last/tests.log 2>&1; echo "exit $?" >> gpurun_out/last/tests.log; tail -15 gpurun_out/last/tests.log
