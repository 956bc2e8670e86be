#!/bin/bash
# round-2 GPU batch 44 (one GPU): host-overhead trims of the eager path: all tests, eager profile, bench c3full + train
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b44_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b44_pytest.log
tail -5 gpurun_out/b44_pytest.log | cut -c1-300
python tools/profile_eager.py > gpurun_out/b44_profile.txt 2>&1; head -14 gpurun_out/b44_profile.txt | cut -c1-150
timeout 900 python bench.py --no-c4 --no-e2e --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/b44_bench.json 2> gpurun_out/b44_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b44_bench.json') if l.startswith('{')][-1]); c=d['c3full']
print("c3full graph", c['ms_per_step'], c['e2e']['ms_per_step'], "eager", c['eager']['ms_per_step'], c['eager']['e2e']['ms_per_step'])
t=d['train_c3']; print("train eager", t['train_step_ms'], "graph", t['train_step_graphed_ms'])
PY
tail -3 gpurun_out/b44_bench.err
