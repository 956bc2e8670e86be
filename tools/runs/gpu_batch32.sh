#!/bin/bash
# round-2 GPU batch 32 (one GPU): the default bench line after the cross-step e2e pipelining (8 chunks, two result buffers)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/b32_bench_n1.json 2> gpurun_out/b32_bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b32_bench_n1.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'])
print({k:v for k,v in d['e2e'].items() if k not in ('path','note')})
c=d['c3full']; print({k:(c[k] if not isinstance(c[k],dict) else {a:b for a,b in c[k].items() if a not in ('path','sample')}) for k in c if k!='workload'})
print({k:v for k,v in d['train_c3'].items() if k!='workload'})
PY
tail -3 gpurun_out/b32_bench_n1.err
