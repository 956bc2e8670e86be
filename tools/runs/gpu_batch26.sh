#!/bin/bash
# round-2 GPU batch 26 (TWO GPUs): sharded + C5 tests after the tower / optimizer changes; sharded bench at B = 8192 with phase breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_shard_kernels_gpu.py -m gpu -q > gpurun_out/b26_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b26_pytest.log
tail -5 gpurun_out/b26_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/bench_sharded.py --batch 8192 > gpurun_out/b26_c4_n2.json 2> gpurun_out/b26_c4_n2.err; echo "bench_sharded exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b26_c4_n2.json') if l.startswith('{')][-1])
print(json.dumps(d.get('train_c5'))[:600]); print(json.dumps(d['summary'])[:900])
print({k:v for k,v in d['ms_per_step'].items() if 'phases' in k})
PY
tail -3 gpurun_out/b26_c4_n2.err
