#!/bin/bash
# round-2 GPU batch 52 (one GPU): evidence -- all tests, smoke, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b52_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b52_pytest.log
tail -4 gpurun_out/b52_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/b52_smoke.log 2>&1; tail -3 gpurun_out/b52_smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/b52_bench_ref.json 2> gpurun_out/b52_bench_ref.err; echo "ref exit $?"
timeout 1500 python bench.py > gpurun_out/b52_bench_n1.json 2> gpurun_out/b52_bench_n1.err; echo "bench exit $?"; tail -c 300 gpurun_out/b52_bench_n1.json; tail -3 gpurun_out/b52_bench_n1.err
