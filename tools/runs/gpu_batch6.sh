#!/bin/bash
# round-2 GPU batch 6 (TWO GPUs): sharded tests incl. the C-ABI step, full bench.py at N = 2 (c2 + e2e + sharded C4), GEMM bench on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_bag_gpu.py tests/test_model_gpu.py -m gpu -x -q > gpurun_out/b6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b6_pytest.log
tail -8 gpurun_out/b6_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b6_bench_n2.json 2> gpurun_out/b6_bench_n2.err; echo "bench n2 exit $?"
tail -c 2500 gpurun_out/b6_bench_n2.json; tail -5 gpurun_out/b6_bench_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/bench_gemm.py > gpurun_out/b6_gemm.json 2> gpurun_out/b6_gemm.err; cat gpurun_out/b6_gemm.json; tail -3 gpurun_out/b6_gemm.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-c4 --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/b6_bench_c3.json 2> gpurun_out/b6_bench_c3.err; tail -c 1500 gpurun_out/b6_bench_c3.json; tail -3 gpurun_out/b6_bench_c3.err
