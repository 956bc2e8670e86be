#!/bin/bash
# round-2 GPU batch 30 (one GPU): CTA-pair (cta_group::2) variant of the tcgen05 Dense kernel: tests, shapes with pairs on / off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/b30_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b30_pytest.log
tail -6 gpurun_out/b30_pytest.log
for pair in 1 0; do
  RF_DENSE_PAIR=$pair timeout 300 python tools/bench_gemm.py --train --steps 10 > gpurun_out/b30_gemm_train_pair$pair.json 2> gpurun_out/b30_err.txt
  RF_DENSE_PAIR=$pair timeout 300 python tools/bench_gemm.py --steps 10 > gpurun_out/b30_gemm_fwd_pair$pair.json 2>> gpurun_out/b30_err.txt
done
python - <<'PY'
import json
for kind in ("train","fwd"):
    a=json.load(open(f"gpurun_out/b30_gemm_{kind}_pair1.json")); b=json.load(open(f"gpurun_out/b30_gemm_{kind}_pair0.json"))
    for k in a: print(kind,k,"pair",round(a[k]["ms"],4),"single",round(b[k]["ms"],4),"lib",round(a[k]["cublas_tf32_matmul_only_ms"],4), round(a[k]["tflops"]),"TF/s")
PY
tail -3 gpurun_out/b30_err.txt
