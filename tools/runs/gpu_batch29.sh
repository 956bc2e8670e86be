#!/bin/bash
# round-2 GPU batch 29 (one GPU): all tests after the kernel-node descriptor upload; c3full (pipelined graph e2e); train step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b29_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b29_pytest.log
tail -4 gpurun_out/b29_pytest.log
timeout 900 python bench.py --no-c4 --no-e2e --steps 5 --warmup 3 > gpurun_out/b29_bench.json 2> gpurun_out/b29_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b29_bench.json') if l.startswith('{')][-1]); c=d['c3full']
print({k:(c[k] if not isinstance(c[k],dict) else {a:b for a,b in c[k].items() if a!='path'}) for k in ('value','ms_per_step','e2e','e2e_pipelined','parity_check') if k in c})
t=d['train_c3']; print({k:t[k] for k in t if k!='workload'})
PY
tail -3 gpurun_out/b29_bench.err
