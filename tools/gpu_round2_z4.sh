#!/bin/bash
# round 2, call Z4: config 5 with the 2160-point column kernel's strips staged by the TMA unit (LHG_COL_TMA=2: that also
# selects the paired adjoint loop there) against the default (cp.async, one barrier per depth in the adjoint)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for t in 1 2 1 2; do
  LHG_COL_TMA=$t timeout 900 python bench.py --workload c5 --steps 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/z4_c5_tma$t.json 2> gpurun_out/z4_c5_tma$t.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/z4_c5_tma$t.json")); print("c5 LHG_COL_TMA=$t", round(d["value"]), round(d["ms_per_step"], 2), {k: round(x["ms_per_step"], 2) for k, x in d["roofline"]["per_kernel"].items()}, d["loss"])
except Exception as e: print("c5 tma$t ERR", e)
PY
done
