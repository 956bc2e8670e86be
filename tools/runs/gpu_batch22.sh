#!/bin/bash
# round-2 GPU batch 22 (one GPU): ncu --set full of the shipped kernels (bag_forward, sdpa_tc, logits_bf16, dense_tc, prep_bf16)
# on a short bench run that has just exited 0 without ncu; the report is summarised on the box (it is larger than what travels back)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export RF_BENCH_C3_STEPS=1 RF_BENCH_C3_WARMUP=1
SHORT="python bench.py --steps 2 --warmup 3 --no-c4 --no-train --no-cpu-baseline --no-e2e"
timeout 600 $SHORT > gpurun_out/b22_short.json 2> gpurun_out/b22_short.err && \
timeout 1500 ncu --set full --clock-control none -k regex:'bag_forward_kernel|sdpa_tc_kernel|logits_bf16_kernel|dense_tc_kernel|prep_bf16_kernel' -c 60 \
  -o /tmp/r2d_shipped $SHORT > gpurun_out/b22_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/b22_ncu.log
python profiles/summarize_ncu.py /tmp/r2d_shipped.ncu-rep gpurun_out/r2d_ncu_full_shipped_kernels_summary.csv > gpurun_out/b22_summary.txt 2>&1
tail -8 gpurun_out/b22_summary.txt; ls -la gpurun_out/
