#!/bin/bash
# round-2 GPU batch 8 (one GPU): dense tests (bf16 logits, 8-warp Dense epilogue), GEMM / logits benches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_model_gpu.py tests/test_training_gpu.py -m gpu -x -q > gpurun_out/b8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b8_pytest.log
tail -12 gpurun_out/b8_pytest.log
timeout 300 python tools/bench_gemm.py > gpurun_out/b8_gemm.json 2> gpurun_out/b8_gemm.err; cat gpurun_out/b8_gemm.json; tail -3 gpurun_out/b8_gemm.err
timeout 300 python tools/bench_logits.py --skip-fp32 --big > gpurun_out/b8_dense.json 2> gpurun_out/b8_dense.err; cat gpurun_out/b8_dense.json; tail -3 gpurun_out/b8_dense.err
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b8_recall.json 2> gpurun_out/b8_recall.err; cat gpurun_out/b8_recall.json; tail -3 gpurun_out/b8_recall.err
