#!/bin/bash
# round-2 GPU batch 34 (one GPU): min / max pooling backward, LookupEmbedding reference_rows flag, all tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b34_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/b34_pytest.log
tail -8 gpurun_out/b34_pytest.log
