#!/bin/bash
# round 2, call Z2: the fused row kernel as clusters of two CTAs whose warps meet before their gathers (LHG_ROWS_PAIR=1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
LHG_ROWS_PAIR=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config4 or full_size or fused or repeated or uint8 or workspace" > gpurun_out/z2_pytest.log 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/z2_pytest.log
for t in 0 1 0 1; do
  LHG_ROWS_PAIR=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/z2_pair$t.json 2>gpurun_out/z2_pair$t.err
  echo "LHG_ROWS_PAIR=$t"; python tools/bsum.py gpurun_out/z2_pair$t.json
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference"
LHG_ROWS_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'row_inv_fwd' -c 6 --csv --log-file gpurun_out/z2_launches.csv $CMD > gpurun_out/z2_ncu1.log 2>&1
echo "ncu rc $?"; grep -o '"row_inv_fwd[^"]*"\|"dram__bytes_read.sum","[A-Za-z]*","[0-9.]*"\|"gpu__time_duration.sum","[a-z]*","[0-9.]*"' gpurun_out/z2_launches.csv | grep -v row_inv | head -8
