#!/bin/bash
# round-2 GPU batch 25 (one GPU): GEMM shapes + train step after the split-K cost fix; then the ncu launch list of bench.py
# (gpu__time_duration only) on a command that has just exited 0 without ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_gpu.py -m gpu -q -k "split_k or tower or multi_head" > gpurun_out/b25_pytest.log 2>&1; tail -2 gpurun_out/b25_pytest.log
timeout 300 python tools/bench_gemm.py --train --steps 10 > gpurun_out/b25_gemm_train.json 2> gpurun_out/b25_err.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/b25_gemm_train.json"))
for k,v in d.items(): print(k,round(v["ms"],4),"lib",round(v["cublas_tf32_matmul_only_ms"],4), round(v["tflops"]),"TF/s")
PY
STEPS=10 timeout 600 python tools/bench_train.py > gpurun_out/b25_train.json 2> gpurun_out/b25_train.err; cat gpurun_out/b25_train.json; tail -2 gpurun_out/b25_train.err
export RF_BENCH_C3_STEPS=2 RF_BENCH_C3_WARMUP=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c4"
timeout 900 $CMD > gpurun_out/b25_short.json 2> gpurun_out/b25_short.err && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d_ncu_launches.csv $CMD > gpurun_out/b25_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2d_ncu_launches.csv; tail -2 gpurun_out/b25_ncu.log | cut -c1-300
