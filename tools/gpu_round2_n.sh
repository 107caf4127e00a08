#!/bin/bash
# round 2, call N: final build -- full -m gpu suite, every N = 1 bench line, ncu launch list + DRAM bytes, ncu --set full
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/n_pytest.log; tail -3 gpurun_out/n_pytest.log
timeout 900 python bench.py > gpurun_out/n_bench_c4.json 2> gpurun_out/n_bench_c4.err; echo "c4 rc $?"
timeout 300 python bench.py --workload c2 > gpurun_out/n_bench_c2.json 2> gpurun_out/n_bench_c2.err; echo "c2 rc $?"
timeout 900 python bench.py --workload c5 --steps 5 > gpurun_out/n_bench_c5.json 2> gpurun_out/n_bench_c5.err; echo "c5 rc $?"
timeout 300 python bench.py --workload c3 --steps 20 > gpurun_out/n_bench_c3.json 2> gpurun_out/n_bench_c3.err; echo "c3 rc $?"
for f in c4 c2 c5 c3; do python tools/bsum.py gpurun_out/n_bench_$f.json 2>/dev/null | cut -c1-300; done
timeout 900 python bench.py --impl reference > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench_ref.err; echo "ref rc $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/n_smoke.log
