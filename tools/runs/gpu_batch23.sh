#!/bin/bash
# round-2 GPU batch 23 (one GPU): column-tile width of the tcgen05 Dense kernel on the training / tower shapes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for bn in 64 128 256; do
  RF_DENSE_BN=$bn timeout 300 python tools/bench_gemm.py --train --steps 10 > gpurun_out/b23_gemm_train_bn$bn.json 2> gpurun_out/b23_err.txt
  RF_DENSE_BN=$bn timeout 300 python tools/bench_gemm.py --steps 10 > gpurun_out/b23_gemm_fwd_bn$bn.json 2>> gpurun_out/b23_err.txt
done
python - <<'PY'
import json
for kind in ("train","fwd"):
    rows={}
    for bn in (64,128,256):
        d=json.load(open(f"gpurun_out/b23_gemm_{kind}_bn{bn}.json"))
        for k,v in d.items(): rows.setdefault(k,{})[bn]=v["ms"]; rows[k]["lib"]=v["cublas_tf32_matmul_only_ms"]
    for k,v in rows.items(): print(kind,k,{a:round(b,4) for a,b in v.items()})
PY
tail -2 gpurun_out/b23_err.txt
