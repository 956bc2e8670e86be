#!/bin/bash
# round-2 GPU batch 53 (EIGHT GPUs): the bench line at N = 8 (C2 replicas, e2e with the box ceiling, sharded C4 all variants with
# parity on every rank and phase breakdowns, train_c5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 > gpurun_out/b53_bench_n8.json 2> gpurun_out/b53_bench_n8.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b53_bench_n8.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
c4=d['sharded_c4']; print({k:(round(v['best_ms'],4), round(v.get('speedup_vs_n1',0),3)) for k,v in c4['summary'].items()})
print({k:v for k,v in c4['ms_per_step'].items() if 'phases' in k}); print(json.dumps(c4['train_c5'])[:500]); print(c4['parity_ok'])
PY
tail -3 gpurun_out/b53_bench_n8.err | cut -c1-300
