#!/bin/bash
# round-2 GPU batch 20 (one GPU): all tests; the GEMM shapes of a training step; dense + train benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b20_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b20_pytest.log
tail -5 gpurun_out/b20_pytest.log
timeout 300 python tools/bench_gemm.py --train > gpurun_out/b20_gemm.json 2> gpurun_out/b20_gemm.err; cat gpurun_out/b20_gemm.json; tail -3 gpurun_out/b20_gemm.err
timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b20_dense.json 2> gpurun_out/b20_dense.err; cat gpurun_out/b20_dense.json; tail -2 gpurun_out/b20_dense.err
STEPS=10 PROFILE=gpurun_out/b20_train_profile.txt timeout 600 python tools/bench_train.py > gpurun_out/b20_train.json 2> gpurun_out/b20_train.err; cat gpurun_out/b20_train.json; tail -3 gpurun_out/b20_train.err
