#!/bin/bash
# round 2, call C: full -m gpu suite (all failures), ncu launch list with DRAM bytes, ncu --set full of the fused row
# kernel and the column kernel (source-level stall samples)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference"
$CMD > gpurun_out/c_plain.json 2> gpurun_out/c_plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'row_|col_' -c 60 --csv --log-file gpurun_out/c_launches.csv $CMD > gpurun_out/c_ncu1.log 2>&1
echo "ncu launches rc $?"
$CMD > gpurun_out/c_plain2.json 2> gpurun_out/c_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:'row_inv_fwd_fused|col_warp' -s 6 -c 3 -f -o gpurun_out/c_prof $CMD > gpurun_out/c_ncu2.log 2>&1
echo "ncu full rc $?"
ls -la gpurun_out/c_prof* 2>/dev/null
