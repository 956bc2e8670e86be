#!/bin/bash
# round-2 GPU batch 36 (one GPU): ncu --set full of the sharded step's route / combine kernels (one-GPU emulation of an 8-way rank)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python tools/emu_sharded.py --batch 65536 --steps 2 --pool-ctas 0"
timeout 900 $CMD > gpurun_out/b36_plain.log 2>&1 && \
timeout 2400 ncu --set full --clock-control none -k regex:'shard_route_tile_kernel|combine_partials_kernel' -c 8 -o /tmp/r2e_shard $CMD > gpurun_out/b36_ncu.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/b36_ncu.log | cut -c1-200
python profiles/summarize_ncu.py /tmp/r2e_shard.ncu-rep gpurun_out/r2e_ncu_full_shard_route_combine_summary.csv > gpurun_out/b36_summary.txt 2>&1
head -c 1200 gpurun_out/b36_summary.txt; ls -la gpurun_out/ | tail -4
