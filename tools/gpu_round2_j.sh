#!/bin/bash
# round 2, call J: which warps skip the radix-18 pass (A/B), parity of the adjoint/forward paths
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or adjoint or gradients or sharded or config4 or config5 or multi_distance" > gpurun_out/j_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/j_pytest.log
bash tools/exp.sh base p0low base p0low 2>&1 | tee gpurun_out/j_exp.log
