#!/bin/bash
# round 2, call T: build of the split-phase column kernel / warp-local rows -- full -m gpu suite, every N = 1 bench line,
# ncu launch list + DRAM bytes, ncu --set full
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
P=${1:-t}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/${P}_pytest.log; tail -3 gpurun_out/${P}_pytest.log
timeout 900 python bench.py > gpurun_out/${P}_bench_c4.json 2> gpurun_out/${P}_bench_c4.err; echo "c4 rc $?"
timeout 300 python bench.py --workload c2 > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err; echo "c2 rc $?"
timeout 900 python bench.py --workload c5 --steps 5 > gpurun_out/${P}_bench_c5.json 2> gpurun_out/${P}_bench_c5.err; echo "c5 rc $?"
timeout 300 python bench.py --workload c3 --steps 20 > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err; echo "c3 rc $?"
for f in c4 c2 c5 c3; do python tools/bsum.py gpurun_out/${P}_bench_$f.json 2>/dev/null | cut -c1-300; done
timeout 900 python bench.py --impl reference > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; echo "ref rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/${P}_smoke.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-reference"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'row_|col_' -c 60 --csv --log-file gpurun_out/${P}_launches.csv $CMD > gpurun_out/${P}_ncu1.log 2>&1
echo "ncu launches rc $?"
ncu --set full --clock-control none --import-source on -k regex:'row_inv_fwd_fused|col_warp' -s 6 -c 3 -f -o gpurun_out/${P}_prof $CMD > gpurun_out/${P}_ncu2.log 2>&1
echo "ncu full rc $?"
