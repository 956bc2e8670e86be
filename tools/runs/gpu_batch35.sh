#!/bin/bash
# round-2 GPU batch 35 (one GPU): ncu --set full of the training-step kernels (first matched launches of tools/bench_train.py)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
STEPS=1 GRAPH=0 timeout 900 python tools/bench_train.py > gpurun_out/b35_plain.log 2>&1 && \
STEPS=1 GRAPH=0 timeout 2400 ncu --set full --clock-control none -k regex:'adam_rows_kernel|adam_dense_kernel|adam_prep_kernel|column_pass_kernel|batchnorm_dx_kernel|sdpa_backward_tiled_kernel|ce_coef_kernel|splitk_sum_kernel' \
  -c 46 -o /tmp/r2e_train python tools/bench_train.py > gpurun_out/b35_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/b35_ncu.log | cut -c1-200
python profiles/summarize_ncu.py /tmp/r2e_train.ncu-rep gpurun_out/r2e_ncu_full_train_kernels_summary.csv > gpurun_out/b35_summary.txt 2>&1
head -c 1200 gpurun_out/b35_summary.txt; ls -la gpurun_out/ | tail -4
