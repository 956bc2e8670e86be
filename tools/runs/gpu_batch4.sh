#!/bin/bash
# round-2 GPU batch 4 (TWO GPUs): all GPU tests (incl. the 2-GPU sharded tests) + the sharded C4 benchmark at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b4_pytest.log
tail -12 gpurun_out/b4_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_sharded.py > gpurun_out/b4_c4_n2.json 2> gpurun_out/b4_c4_n2.err; echo "bench_sharded exit $?"
tail -c 6000 gpurun_out/b4_c4_n2.json; tail -5 gpurun_out/b4_c4_n2.err
