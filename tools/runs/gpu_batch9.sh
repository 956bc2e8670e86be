#!/bin/bash
# round-2 GPU batch 9 (one GPU): full GPU test suite, smoke, dense benches, full default bench.py line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b9_pytest.log
tail -6 gpurun_out/b9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b9_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/b9_smoke.log; tail -5 gpurun_out/b9_smoke.log
timeout 300 python tools/bench_logits.py --skip-fp32 > gpurun_out/b9_dense.json 2> gpurun_out/b9_dense.err; cat gpurun_out/b9_dense.json
STEPS=20 timeout 300 python tools/bench_recall.py > gpurun_out/b9_recall.json 2> gpurun_out/b9_recall.err; cat gpurun_out/b9_recall.json
timeout 900 python bench.py > gpurun_out/b9_bench.json 2> gpurun_out/b9_bench.err; echo "bench exit $?"; tail -c 2500 gpurun_out/b9_bench.json; tail -3 gpurun_out/b9_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/b9_bench_ref.json 2> gpurun_out/b9_bench_ref.err; echo "ref exit $?"; cut -c1-400 gpurun_out/b9_bench_ref.json
