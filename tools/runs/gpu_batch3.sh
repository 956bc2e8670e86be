#!/bin/bash
# round-2 GPU batch 3 (one GPU): tests incl. the tcgen05 Dense kernel + SDPA with 8 softmax warps; dense benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b3_pytest.log
tail -15 gpurun_out/b3_pytest.log
timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b3_dense.json 2> gpurun_out/b3_dense.err; cat gpurun_out/b3_dense.json
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b3_recall.json 2> gpurun_out/b3_recall.err; cat gpurun_out/b3_recall.json; tail -3 gpurun_out/b3_recall.err
