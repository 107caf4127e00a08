#!/bin/bash
# round 2, call G: TMA strip staging in the column kernel: parity, then A/B against cp.async / __ldg staging
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/g_pytest.log
LHG_COL_TMA=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or adjoint or gradients or sharded or bitwise or config5 or uint8 or multi_distance or chunking" > gpurun_out/g_pytest2.log 2>&1; echo "pytest(TMA=2) rc $?"; tail -3 gpurun_out/g_pytest2.log
for v in 1 0 1 0; do
  LHG_COL_TMA=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/g_tma$v.json 2> gpurun_out/g_tma$v.err
  echo "LHG_COL_TMA=$v"; python tools/bsum.py gpurun_out/g_tma$v.json
done
for v in 2 0 2 0; do
LHG_COL_TMA=$v timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/g_c5_tma$v.json 2> gpurun_out/g_c5_tma$v.err; echo "c5 LHG_COL_TMA=$v"; python tools/bsum.py gpurun_out/g_c5_tma$v.json
done
