#!/bin/bash
# round-2 GPU batch 46 (one GPU): eager training step after the host-side fixes: tests, profile, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_model_gpu.py tests/test_layers_gpu.py -m gpu -q > gpurun_out/b46_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b46_pytest.log
tail -3 gpurun_out/b46_pytest.log | cut -c1-300
python tools/profile_eager_train.py > gpurun_out/b46_profile.txt 2>&1; head -30 gpurun_out/b46_profile.txt | cut -c1-150
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b46_train.json 2> gpurun_out/b46_train.err; python -c "
import json; d=json.load(open('gpurun_out/b46_train.json')); print({k:v for k,v in d.items() if k!='workload'})"
