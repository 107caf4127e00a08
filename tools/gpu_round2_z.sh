#!/bin/bash
# round 2, call Z: L2 promotion of the row kernels' TMA gathers (LHG_ROWS_TMA_PROMOTE=1 = 128-byte promotion as before)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for t in 1 0 1 0; do
  LHG_ROWS_TMA_PROMOTE=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/z_promo$t.json 2>gpurun_out/z_promo$t.err
  echo "LHG_ROWS_TMA_PROMOTE=$t"; python tools/bsum.py gpurun_out/z_promo$t.json
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'row_|col_' -c 30 --csv --log-file gpurun_out/z_launches.csv $CMD > gpurun_out/z_ncu1.log 2>&1
echo "ncu rc $?"
python tools/ncu_traffic.py gpurun_out/z_launches.csv c4 gpurun_out/z_traffic.json "z" | tail -5
