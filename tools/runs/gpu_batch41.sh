#!/bin/bash
# round-2 GPU batch 41 (one GPU): randomised stress of the Dense kernel (auto, forced pairs, pairs off)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/stress_gemm.py --n 150 --seed 1 > gpurun_out/b41_auto.json 2> gpurun_out/b41_err.txt; echo "auto exit $?"; cat gpurun_out/b41_auto.json | cut -c1-600
RF_DENSE_PAIR=2 timeout 600 python tools/stress_gemm.py --n 150 --seed 2 > gpurun_out/b41_pair.json 2>> gpurun_out/b41_err.txt; echo "pair exit $?"; cat gpurun_out/b41_pair.json | cut -c1-600
RF_DENSE_PAIR=0 timeout 600 python tools/stress_gemm.py --n 100 --seed 3 > gpurun_out/b41_single.json 2>> gpurun_out/b41_err.txt; echo "single exit $?"; cat gpurun_out/b41_single.json | cut -c1-600
tail -3 gpurun_out/b41_err.txt
