#!/bin/bash
# round 2, call A: the full -m gpu suite, then every bench workload once on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt
free -g > gpurun_out/a_host.txt; nproc >> gpurun_out/a_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -30 gpurun_out/a_pytest.log
timeout 120 tools/ubench_tc > gpurun_out/a_ubench_tc.log 2>&1; echo "ubench_tc rc $?"; cat gpurun_out/a_ubench_tc.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_c4.json 2> gpurun_out/a_bench_c4.err; echo "c4 rc $?"
timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err; echo "c2 rc $?"
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/a_bench_c5.json 2> gpurun_out/a_bench_c5.err; echo "c5 rc $?"
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 > gpurun_out/a_bench_c3.json 2> gpurun_out/a_bench_c3.err; echo "c3 rc $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc $?"
for f in c4 c2 c5 c3 ref; do echo "== $f"; head -c 3000 gpurun_out/a_bench_$f.json; tail -5 gpurun_out/a_bench_$f.err; done
