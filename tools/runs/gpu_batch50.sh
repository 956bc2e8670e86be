#!/bin/bash
# round-2 GPU batch 50 (one GPU): 64 x 64 / float4 coefficient pass of the CE backward: tests, train bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_training_gpu.py -m gpu -q > gpurun_out/b50_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b50_pytest.log
tail -3 gpurun_out/b50_pytest.log | cut -c1-300
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b50_train.json 2> gpurun_out/b50_train.err; python -c "
import json; d=json.load(open('gpurun_out/b50_train.json')); print({k:v for k,v in d.items() if k!='workload'})"
