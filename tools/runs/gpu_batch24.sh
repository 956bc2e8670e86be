#!/bin/bash
# round-2 GPU batch 24 (one GPU): model-based tile / split-K choice of the tcgen05 Dense kernel: tests, shapes, recall forward, train step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_model_gpu.py tests/test_training_gpu.py -m gpu -q > gpurun_out/b24_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b24_pytest.log
tail -4 gpurun_out/b24_pytest.log
timeout 300 python tools/bench_gemm.py --train --steps 10 > gpurun_out/b24_gemm_train.json 2> gpurun_out/b24_err.txt
timeout 300 python tools/bench_gemm.py --steps 10 > gpurun_out/b24_gemm_fwd.json 2>> gpurun_out/b24_err.txt
python - <<'PY'
import json
for kind in ("train","fwd"):
    d=json.load(open(f"gpurun_out/b24_gemm_{kind}.json"))
    for k,v in d.items(): print(kind,k,round(v["ms"],4),"lib",round(v["cublas_tf32_matmul_only_ms"],4), round(v["tflops"]),"TF/s")
PY
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b24_recall.json 2> gpurun_out/b24_recall.err; cat gpurun_out/b24_recall.json
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b24_train.json 2> gpurun_out/b24_train.err; cat gpurun_out/b24_train.json; tail -3 gpurun_out/b24_train.err
