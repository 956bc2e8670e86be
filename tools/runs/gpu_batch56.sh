#!/bin/bash
# round-2 GPU batch 56 (FOUR GPUs): the full bench line at N = 4 (the driver's scaling run launches N = 1, 2, 4, 8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 > gpurun_out/b56_bench_n4.json 2> gpurun_out/b56_bench_n4.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b56_bench_n4.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
c4=d['sharded_c4']; print({k:(round(v['best_ms'],4), round(v.get('speedup_vs_n1',0),3)) for k,v in c4['summary'].items()})
print(json.dumps(c4['train_c5'])[:500]); print(c4['parity_ok'])
PY
tail -2 gpurun_out/b56_bench_n4.err | cut -c1-300
