#!/bin/bash
# round-2 GPU batch 1 (one GPU): tests, smoke, SDPA A/B, sharded-step emulation A/B, route kernel profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/b1_clocks.csv &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b1_pytest.log
tail -5 gpurun_out/b1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b1_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/b1_smoke.log
tail -5 gpurun_out/b1_smoke.log
timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b1_dense_v2.json 2> gpurun_out/b1_dense_v2.err
RF_SDPA_V1=1 timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b1_dense_v1.json 2> gpurun_out/b1_dense_v1.err
cat gpurun_out/b1_dense_v2.json gpurun_out/b1_dense_v1.json
for st in 1 0; do
  RF_ROUTE_STAGE=$st timeout 600 python tools/emu_sharded.py --label "stage$st" > gpurun_out/b1_emu_str_stage$st.json 2> gpurun_out/b1_emu_str_stage$st.err
  cat gpurun_out/b1_emu_str_stage$st.json
done
timeout 600 python tools/emu_sharded.py --prehashed --label prehashed > gpurun_out/b1_emu_pre.json 2> gpurun_out/b1_emu_pre.err
cat gpurun_out/b1_emu_pre.json
timeout 600 python tools/emu_sharded.py --batch 8192 --label b8192 > gpurun_out/b1_emu_b8192.json 2> gpurun_out/b1_emu_b8192.err
cat gpurun_out/b1_emu_b8192.json
kill $SMI
# ncu: the route kernel (staged), after the same command exited 0 above without ncu
timeout 300 python tools/route_microbench.py --steps 3 > gpurun_out/b1_route_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shard_route_tile -s 2 -c 2 -o gpurun_out/r2a_route_tile \
    python tools/route_microbench.py --steps 3 > gpurun_out/b1_ncu_route.log 2>&1
tail -3 gpurun_out/b1_ncu_route.log
