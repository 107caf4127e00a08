#!/bin/bash
# round 2, call X: the fused row kernel's per-warp bulk copies (LHG_ROWS_TMA=1) -- parity subset with them on, then A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/x_pytest.log
for t in 0 1 0 1; do
  LHG_ROWS_TMA=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/x_tma$t.json 2>gpurun_out/x_tma$t.err
  echo "LHG_ROWS_TMA=$t"; python tools/bsum.py gpurun_out/x_tma$t.json
done
for t in 0 1; do
  LHG_ROWS_TMA=$t timeout 900 python bench.py --workload c5 --steps 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/x_c5_tma$t.json 2> gpurun_out/x_c5_tma$t.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/x_c5_tma$t.json")); print("c5 tma$t", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["ms_per_step"], 2) for k, x in d["roofline"]["per_kernel"].items()})
except Exception as e: print("c5 tma$t ERR", e)
PY
done
