#!/bin/bash
# round-2 GPU batch 33 (one GPU): e2e stream count sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for s in 2 3 4; do
RF_E2E_STREAMS=$s timeout 600 python bench.py --no-c4 --no-c3 --no-train --no-cpu-baseline > gpurun_out/b33_s$s.json 2> gpurun_out/b33_s$s.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/b33_s$s.json') if l.startswith('{')][-1])
print($s, d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['fraction_of_box_d2h_ceiling'])
PY
done
