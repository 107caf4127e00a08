#!/bin/bash
# round 2, call Q: ncu launch list + DRAM bytes and ncu --set full (source-level stalls) of the split-phase / warp-local build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/q_plain.json 2> gpurun_out/q_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'row_|col_' -c 60 --csv --log-file gpurun_out/q_launches.csv $CMD > gpurun_out/q_ncu1.log 2>&1
echo "ncu launches rc $?"
ncu --set full --clock-control none --import-source on -k regex:'row_inv_fwd_fused|col_warp' -s 6 -c 3 -f -o gpurun_out/q_prof $CMD > gpurun_out/q_ncu2.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out/q_prof.ncu-rep
