#!/bin/bash
# round-2 GPU batch 37 (TWO GPUs): sharded tests incl. the p2p backward and the C5 trainer on both transports; train_c5 at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_sharded_gpu.py -m gpu -q > gpurun_out/b37_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b37_pytest.log
tail -12 gpurun_out/b37_pytest.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 tools/bench_sharded.py --batch 8192 > gpurun_out/b37_c4_n2.json 2> gpurun_out/b37_c4_n2.err; echo "bench_sharded exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b37_c4_n2.json') if l.startswith('{')][-1])
print(json.dumps(d.get('train_c5'))[:900])
PY
tail -3 gpurun_out/b37_c4_n2.err | cut -c1-300
