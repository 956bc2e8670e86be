#!/bin/bash
# round-2 GPU batch 13 (one GPU): pipelined-softmax SDPA + tensor-core CE backward: tests, dense + training benches, step profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b13_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b13_pytest.log
tail -8 gpurun_out/b13_pytest.log
timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b13_dense.json 2> gpurun_out/b13_dense.err; cat gpurun_out/b13_dense.json; tail -2 gpurun_out/b13_dense.err
STEPS=10 PROFILE=gpurun_out/b13_train_profile.txt timeout 600 python tools/bench_train.py > gpurun_out/b13_train.json 2> gpurun_out/b13_train.err; cat gpurun_out/b13_train.json; tail -3 gpurun_out/b13_train.err
