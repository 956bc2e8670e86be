#!/bin/bash
# round-2 GPU batch 7 (TWO GPUs): C5 trainer test, sharded bench with train_c5; on one GPU: c3full with the packed fast path, GEMM profile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_model_gpu.py tests/test_training_gpu.py -m gpu -x -q > gpurun_out/b7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b7_pytest.log
tail -8 gpurun_out/b7_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_sharded.py --batch 8192 > gpurun_out/b7_c4_n2.json 2> gpurun_out/b7_c4_n2.err; echo "bench_sharded exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b7_c4_n2.json') if l.startswith('{')][-1])
print(json.dumps(d.get('train_c5'))); print(d['summary'])
PY
tail -3 gpurun_out/b7_c4_n2.err
export CUDA_VISIBLE_DEVICES=0
timeout 600 python bench.py --no-c4 --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/b7_bench_c3.json 2> gpurun_out/b7_bench_c3.err; tail -c 900 gpurun_out/b7_bench_c3.json; tail -3 gpurun_out/b7_bench_c3.err
timeout 300 python tools/bench_gemm.py --steps 5 > gpurun_out/b7_gemm.json 2> gpurun_out/b7_gemm.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 8 -c 1 -o gpurun_out/r2b_gemm_1888x1024 python tools/bench_gemm.py --steps 5 > gpurun_out/b7_ncu_gemm.log 2>&1
tail -2 gpurun_out/b7_ncu_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -s 84 -c 1 -o gpurun_out/r2b_gemm_64x64 python tools/bench_gemm.py --steps 5 > gpurun_out/b7_ncu_gemm2.log 2>&1
tail -2 gpurun_out/b7_ncu_gemm2.log
