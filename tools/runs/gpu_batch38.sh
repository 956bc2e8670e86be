#!/bin/bash
# round-2 GPU batch 38 (one GPU): tensor-core SDPA backward: tests, train bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b38_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b38_pytest.log
tail -8 gpurun_out/b38_pytest.log | cut -c1-300
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b38_train.json 2> gpurun_out/b38_train.err; python -c "
import json; d=json.load(open('gpurun_out/b38_train.json')); print({k:v for k,v in d.items() if k!='workload'})"; tail -2 gpurun_out/b38_train.err
