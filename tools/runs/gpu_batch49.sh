#!/bin/bash
# round-2 GPU batch 49 (one GPU): sequence encoder overlapped with the bag launch: model tests, c3full
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_training_gpu.py tests/test_graphs_gpu.py -m gpu -q > gpurun_out/b49_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b49_pytest.log
tail -3 gpurun_out/b49_pytest.log | cut -c1-300
timeout 900 python bench.py --no-c4 --no-e2e --no-train --steps 5 --warmup 3 > gpurun_out/b49_bench.json 2> gpurun_out/b49_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b49_bench.json') if l.startswith('{')][-1]); c=d['c3full']
print("c3full graph", c['ms_per_step'], c['e2e']['ms_per_step'], "eager", c['eager']['ms_per_step'], c['eager']['e2e']['ms_per_step'], c['parity_check'])
PY
tail -3 gpurun_out/b49_bench.err
