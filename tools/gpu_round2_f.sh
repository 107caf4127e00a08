#!/bin/bash
# round 2, call F: full -m gpu suite + the default bench line (what the driver runs at N = 1) + smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/f_pytest.log
tail -5 gpurun_out/f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc $?"; tail -3 gpurun_out/f_smoke.log
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?"; head -c 600 gpurun_out/f_bench.json; echo
timeout 900 python bench.py --impl reference > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "ref rc $?"; head -c 400 gpurun_out/f_bench_ref.json; echo
