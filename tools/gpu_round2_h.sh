#!/bin/bash
# round 2, call H: A/B of the adjoint column launch with the depth-sum accumulator in registers; new loss tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_next_stages.py -m gpu -x -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/h_pytest.log
LHG_LIB=$PWD/learned_hologram_gan_b200/lib/libasm_b200_regacc.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or adjoint or gradients or uint8" > gpurun_out/h_pytest2.log 2>&1; echo "pytest(regacc) rc $?"; tail -3 gpurun_out/h_pytest2.log
bash tools/exp.sh base regacc base regacc 2>&1 | tee gpurun_out/h_exp.log
