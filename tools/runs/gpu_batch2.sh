#!/bin/bash
# round-2 GPU batch 2 (one GPU): SDPA v2 profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/bench_logits.py --skip-fp32 --steps 5 > gpurun_out/b2_dense.json 2> gpurun_out/b2_dense.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sdpa_tc_kernel -s 3 -c 1 -o gpurun_out/r2a_sdpa_v2 \
    python tools/bench_logits.py --skip-fp32 --steps 5 > gpurun_out/b2_ncu_sdpa.log 2>&1
tail -3 gpurun_out/b2_ncu_sdpa.log
cat gpurun_out/b2_dense.json
