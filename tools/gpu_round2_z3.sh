#!/bin/bash
# round 2, call Z3: cluster size of the fused row kernel's rendezvous (rows per cluster: 2, 4, 8)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for t in 2 4 8 2 4 8; do
  LHG_ROWS_PAIR=$t timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/z3_pair$t.json 2>gpurun_out/z3_pair$t.err
  echo "LHG_ROWS_PAIR=$t"; python tools/bsum.py gpurun_out/z3_pair$t.json
done
for t in 0 4; do
  LHG_ROWS_PAIR=$t timeout 900 python bench.py --workload c5 --steps 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/z3_c5_pair$t.json 2> gpurun_out/z3_c5_pair$t.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/z3_c5_pair$t.json")); print("c5 pair$t", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["ms_per_step"], 2) for k, x in d["roofline"]["per_kernel"].items()})
except Exception as e: print("c5 pair$t ERR", e)
PY
done
