#!/bin/bash
# round 2, call B: A/B of the 3-pass row plan, then the full -m gpu suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
bash tools/exp.sh base rows4 minb2 2>&1 | tee gpurun_out/b_exp.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/b_pytest.log
tail -40 gpurun_out/b_pytest.log
