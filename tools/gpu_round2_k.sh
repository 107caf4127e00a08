#!/bin/bash
# round 2, call K: ncu --set full of the C2 (1024 x 1024) column and fused row kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/k_plain.json 2> gpurun_out/k_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'row_inv_fwd_fused|col_warp16' -s 6 -c 3 -f -o gpurun_out/k_prof $CMD > gpurun_out/k_ncu.log 2>&1
echo "ncu full rc $?"
