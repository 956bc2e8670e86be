#!/bin/bash
# round-2 GPU batch 48 (TWO GPUs): the full bench line at N = 2 (what the driver's scaling run launches) + the reference arm under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/b48_bench_n2.json 2> gpurun_out/b48_bench_n2.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b48_bench_n2.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
c4=d['sharded_c4']; print({k:(round(v['best_ms'],4), round(v.get('speedup_vs_n1',0),3)) for k,v in c4['summary'].items()})
print(json.dumps(c4['train_c5'])[:400]); print(c4['parity_ok'])
PY
tail -3 gpurun_out/b48_bench_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-200
