#!/bin/bash
# round-2 GPU batch 31 (one GPU): gated CTA pairs: all tests, GEMM shapes (auto), recall forward, train step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b31_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b31_pytest.log
tail -6 gpurun_out/b31_pytest.log
timeout 300 python tools/bench_gemm.py --train --steps 10 > gpurun_out/b31_gemm_train.json 2> gpurun_out/b31_err.txt
timeout 300 python tools/bench_gemm.py --steps 10 > gpurun_out/b31_gemm_fwd.json 2>> gpurun_out/b31_err.txt
python - <<'PY'
import json
for kind in ("train","fwd"):
    a=json.load(open(f"gpurun_out/b31_gemm_{kind}.json"))
    for k in a: print(kind,k,round(a[k]["ms"],4),"lib",round(a[k]["cublas_tf32_matmul_only_ms"],4), round(a[k]["tflops"]),"TF/s")
PY
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b31_recall.json 2> gpurun_out/b31_recall.err; cat gpurun_out/b31_recall.json
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b31_train.json 2> gpurun_out/b31_train.err; python -c "
import json; d=json.load(open('gpurun_out/b31_train.json')); print({k:v for k,v in d.items() if k!='workload'})"
