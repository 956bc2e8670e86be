#!/bin/bash
# round-2 GPU batch 16 (one GPU): training stage on tcgen05, graphed train step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b16_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b16_pytest.log
tail -12 gpurun_out/b16_pytest.log
STEPS=10 GRAPH=1 PROFILE=gpurun_out/b16_train_profile.txt timeout 600 python tools/bench_train.py > gpurun_out/b16_train.json 2> gpurun_out/b16_train.err; cat gpurun_out/b16_train.json; tail -5 gpurun_out/b16_train.err
