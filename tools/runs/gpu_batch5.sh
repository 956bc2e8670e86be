#!/bin/bash
# round-2 GPU batch 5 (one GPU): route/pool co-residency variants in the one-GPU emulation; tests of the new pieces
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_shard_kernels_gpu.py tests/test_dense_gpu.py tests/test_training_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/b5_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b5_pytest.log
tail -6 gpurun_out/b5_pytest.log
for cap in 0 1 2; do
  RF_ROUTE_CTAS_PER_SM=$cap timeout 600 python tools/emu_sharded.py --label "routecap$cap" --pool-ctas 0 3 2 > gpurun_out/b5_emu_cap$cap.json 2> gpurun_out/b5_emu_cap$cap.err
  cat gpurun_out/b5_emu_cap$cap.json
done
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b5_recall.json 2> gpurun_out/b5_recall.err; cat gpurun_out/b5_recall.json; tail -3 gpurun_out/b5_recall.err
timeout 600 python bench.py --no-c4 --steps 10 --warmup 3 > gpurun_out/b5_bench.json 2> gpurun_out/b5_bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/b5_bench.json; tail -5 gpurun_out/b5_bench.err
