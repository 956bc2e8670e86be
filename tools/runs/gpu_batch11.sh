#!/bin/bash
# round-2 GPU batch 11 (one GPU): tests after the SDPA / training / encoder changes; dense + training benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b11_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b11_pytest.log
tail -8 gpurun_out/b11_pytest.log
timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b11_dense.json 2> gpurun_out/b11_dense.err; cat gpurun_out/b11_dense.json; tail -2 gpurun_out/b11_dense.err
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b11_train.json 2> gpurun_out/b11_train.err; cat gpurun_out/b11_train.json; tail -3 gpurun_out/b11_train.err
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b11_recall.json 2> gpurun_out/b11_recall.err; cat gpurun_out/b11_recall.json
