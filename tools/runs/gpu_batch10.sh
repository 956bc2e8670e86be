#!/bin/bash
# round-2 GPU batch 10 (EIGHT GPUs): the default bench.py line at N = 8 (c2 replicas + e2e + PCIe ceiling + sharded C4 with oracle parity + train_c5)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/b10_bench_n8.json 2> gpurun_out/b10_bench_n8.err; echo "bench n8 exit $?"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/b10_bench_n8.json') if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print(d['e2e'])
    c=d['sharded_c4']; print(json.dumps(c['summary'])); print(c['parity_ok'], c['limiter']); print(c['train_c5'])
    for k,v in c['ms_per_step'].items(): print(k,v)
except Exception as e: print('parse failed', e)
PY
tail -5 gpurun_out/b10_bench_n8.err
