#!/bin/bash
# round 2, call R: A/B of the L2 prefetch of the column kernel's strips (variant nopf = without) + parity subset
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config4 or full_size or fused or sharded or repeated or config5 or multi_distance or forward_methods or gradients" > gpurun_out/r_pytest.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r_pytest.log
bash tools/exp.sh "$@"
