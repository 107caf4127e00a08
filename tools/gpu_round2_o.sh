#!/bin/bash
# round 2, call O: split-phase (mbarrier) depth loop of the column kernel -- parity on the variant library, then A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
V=${1:-split}
LHG_LIB=$PWD/learned_hologram_gan_b200/lib/libasm_b200_$V.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "config4 or full_size or fused or multi_distance or repeated or config5 or sharded" > gpurun_out/o_pytest.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/o_pytest.log
bash tools/exp.sh base split $V base split $V
