#!/bin/bash
# round-2 GPU batch 42 (one GPU): compute-sanitizer memcheck over the tests of the kernels added this round
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SEL="pairs or split_k or column_passes or training_stage or sdpa_backward or minmax or live_row or fused_qkv or graphed_call or multi_table"
timeout 600 python -m pytest tests/test_dense_gpu.py tests/test_bag_gpu.py tests/test_graphs_gpu.py -m gpu -q -k "$SEL" > gpurun_out/b42_plain.log 2>&1; echo "plain exit $?"; tail -2 gpurun_out/b42_plain.log
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 30 python -m pytest tests/test_dense_gpu.py tests/test_bag_gpu.py tests/test_graphs_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/b42_memcheck.log 2>&1; echo "memcheck exit $?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/b42_memcheck.log; grep -m5 -B2 -A12 "Invalid\|misaligned" gpurun_out/b42_memcheck.log | cut -c1-220 | head -60; tail -6 gpurun_out/b42_memcheck.log | cut -c1-200
