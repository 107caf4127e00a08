#!/bin/bash
# round 2, call M: ncu --set full of the C5 (1080p, padded) column and fused row kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/m_plain.json 2> gpurun_out/m_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'row_inv_fwd_fused|col_warp' -s 4 -c 3 -f -o gpurun_out/m_prof $CMD > gpurun_out/m_ncu.log 2>&1
echo "ncu full rc $?"
