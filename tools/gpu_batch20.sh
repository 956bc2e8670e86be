#!/bin/bash
# round-2 GPU batch 20 (one GPU): the GEMM shapes of a training step, dense tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_gpu.py -m gpu -q > gpurun_out/b20_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b20_pytest.log
tail -5 gpurun_out/b20_pytest.log
timeout 300 python tools/bench_gemm.py --train > gpurun_out/b20_gemm.json 2> gpurun_out/b20_gemm.err; cat gpurun_out/b20_gemm.json; tail -3 gpurun_out/b20_gemm.err
