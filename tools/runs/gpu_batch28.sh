#!/bin/bash
# round-2 GPU batch 28 (one GPU): one-GPU emulation of an 8-way rank at B = 8192 (is the slow small-batch pool the kernel or NVLink?),
# c3full with the pipelined graph e2e, smoke with the Dense kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/emu_sharded.py --batch 8192 --label b8192 > gpurun_out/b28_emu_b8192.json 2> gpurun_out/b28_emu.err; cat gpurun_out/b28_emu_b8192.json; tail -2 gpurun_out/b28_emu.err
timeout 600 python tools/emu_sharded.py --batch 65536 --label b65536 > gpurun_out/b28_emu_b65536.json 2>> gpurun_out/b28_emu.err; cat gpurun_out/b28_emu_b65536.json
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b28_smoke.log 2>&1; tail -2 gpurun_out/b28_smoke.log
timeout 900 python bench.py --no-c4 --no-train --no-e2e --steps 5 --warmup 3 > gpurun_out/b28_bench.json 2> gpurun_out/b28_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/b28_bench.json') if l.startswith('{')][-1]); c=d['c3full']
print({k:c[k] for k in ('value','ms_per_step','e2e','e2e_pipelined','parity_check') if k in c})
PY
tail -3 gpurun_out/b28_bench.err
