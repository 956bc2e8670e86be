#!/bin/bash
# round-2 GPU batch 55 (one GPU): ncu --set full of the Dense kernel on the tower shapes (CTA pairs on the first two)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python tools/bench_gemm.py --steps 2"
timeout 600 $CMD > gpurun_out/b55_plain.json 2> gpurun_out/b55_plain.err && \
timeout 1500 ncu --set full --clock-control none -k regex:dense_tc_kernel -c 15 -o /tmp/r2g_dense $CMD > gpurun_out/b55_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/b55_ncu.log | cut -c1-200
python profiles/summarize_ncu.py /tmp/r2g_dense.ncu-rep gpurun_out/r2g_ncu_full_dense_tc_pairs_summary.csv > gpurun_out/b55_summary.txt 2>&1
head -c 1500 gpurun_out/b55_summary.txt
