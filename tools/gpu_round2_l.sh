#!/bin/bash
# round 2, call L: bank-conflict swizzle of the 1024-point rows: full GPU suite, C2 / C3 lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/l_pytest.log
for i in 1 2; do
timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-gpu-reference > gpurun_out/l_c2.json 2> gpurun_out/l_c2.err; python tools/bsum.py gpurun_out/l_c2.json
done
timeout 300 python bench.py --workload c3 --steps 20 > gpurun_out/l_c3.json 2> gpurun_out/l_c3.err; python -c "
import json; d=json.load(open('gpurun_out/l_c3.json')); print(d['value'], d['step'])"
